set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu10.log
python - <<'PY' 2>&1 | tee gpurun_out/render_bench.log
import torch, time, sys
sys.path.insert(0, '.')
from bench import baked_triplane
from sculptmate_b200.tsr import TSR
torch.manual_seed(0)
m = TSR().cuda()
tp = baked_triplane(100).cuda()
H = W = 512
g = torch.Generator().manual_seed(1)
o = torch.randn(H * W, 3, generator=g); o = (o / o.norm(dim=-1, keepdim=True) * 2.0).cuda()
t = ((torch.rand(H * W, 3, generator=g) * 2 - 1) * 0.6).cuda()
d = t - o; d = d / d.norm(dim=-1, keepdim=True)
for prec in ("tc", "fp32"):
    for i in range(3):
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); c = m.renderer(m.decoder, tp, o.view(H, W, 3), d.view(H, W, 3), precision=prec); b.record(); torch.cuda.synchronize()
    print(f"render 512x512 rays x 128 samples ({H*W*128/1e6:.1f} M samples) precision={prec}: {a.elapsed_time(b):.2f} ms, mean rgb {float(c.mean()):.4f}")
PY
