set -x
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu12.log
python - <<'PY' 2>&1 | tee gpurun_out/bake_bench.log
import sys, numpy as np, torch
sys.path.insert(0, '.')
from sculptmate_b200.sf3d import TextureBaker
rng = np.random.RandomState(0)
n = 224  # ~100 k triangles
gx = np.linspace(0.0, 1.0, n)
u, v = np.meshgrid(gx, gx, indexing="ij")
uv = (np.stack([u, v], -1).reshape(-1, 2) + rng.uniform(-0.001, 0.001, (n * n, 2))).astype(np.float32)
idx = np.arange(n * n).reshape(n, n)
a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
faces = np.concatenate([np.stack([a, b, c], -1), np.stack([a, c, d], -1)]).astype(np.int32)
attr = rng.randn(n * n, 3).astype(np.float32)
tb = TextureBaker()
uvd, fd, ad = torch.from_numpy(uv).cuda(), torch.from_numpy(faces).cuda(), torch.from_numpy(attr).cuda()
for res in (1024, 2048):
    for i in range(3):
        torch.cuda.synchronize(); e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); r = tb.rasterize(uvd, fd, res, "cuda"); e[1].record(); o = tb.interpolate(ad, r, fd, res, "cuda"); e[2].record(); torch.cuda.synchronize()
    print(f"bake res {res}, {len(faces)} triangles: rasterize {e[0].elapsed_time(e[1]):.3f} ms, interpolate {e[1].elapsed_time(e[2]):.3f} ms, coverage {float((r[...,3]>=0).float().mean()):.3f}")
PY
