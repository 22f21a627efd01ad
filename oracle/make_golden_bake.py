#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- generates tests/golden/bake.npz by running the reference's own Python functions
rasterize_cpu / interpolate_cpu (/root/reference/StableFast/sf3d/texture_baker/common.py:123-142, 214-230; the
functions the production path calls by the same names inside texture_baker.dll, baker.py:30-57,92-118) on a small
UV atlas: a jittered triangulated grid, a separate island, unused space, one degenerate triangle.

    python oracle/make_golden_bake.py        (needs /root/reference; numpy >= 2 semantics: fp32 scalars)
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("SCULPTMATE_REFERENCE_ROOT", "/root/reference")


def load_common():
    spec = importlib.util.spec_from_file_location("ref_texture_baker_common", os.path.join(REF, "StableFast", "sf3d", "texture_baker", "common.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def atlas(seed: int, n: int = 7):
    rng = np.random.RandomState(seed)
    g = np.linspace(0.08, 0.8, n)
    u, v = np.meshgrid(g, g, indexing="ij")
    uv = np.stack([u, v], -1).reshape(-1, 2)
    uv += rng.uniform(-0.02, 0.02, uv.shape)  # jitter smaller than half a cell: no fold-overs
    faces = []
    for i in range(n - 1):
        for j in range(n - 1):
            a, b, c, d = i * n + j, (i + 1) * n + j, (i + 1) * n + j + 1, i * n + j + 1
            faces += [[a, b, c], [a, c, d]] if (i + j) % 2 == 0 else [[a, b, d], [b, c, d]]
    base = len(uv)
    island = np.array([[0.86, 0.86], [0.99, 0.87], [0.93, 0.99], [0.87, 0.97]])
    uv = np.concatenate([uv, island, [[0.5, 0.95], [0.5, 0.95], [0.6, 0.95]]], 0)
    faces += [[base, base + 1, base + 2], [base, base + 2, base + 3]]
    faces += [[base + 4, base + 5, base + 6]]  # degenerate (two identical vertices)
    attr = rng.randn(len(uv), 3)
    return uv.astype(np.float32), np.asarray(faces, np.int32), attr.astype(np.float32)


def main() -> None:
    tb = load_common()
    out = {}
    for k, (seed, res) in enumerate([(0, 48), (1, 64)]):
        uv, faces, attr = atlas(seed)
        with np.errstate(all="ignore"):
            rast = tb.rasterize_cpu(uv, faces, res)
            interp = tb.interpolate_cpu(attr, faces, rast)
        cov = float((rast[..., 3] >= 0).mean())
        print(f"case {k}: res {res}, {len(faces)} faces, coverage {cov:.3f}")
        assert 0.3 < cov < 0.9
        out.update({f"uv{k}": uv, f"faces{k}": faces, f"attr{k}": attr, f"rast{k}": rast.astype(np.float32), f"interp{k}": interp.astype(np.float32),
                    f"res{k}": np.int64(res)})
    np.savez_compressed(os.path.join(GOLD, "bake.npz"), **out)


if __name__ == "__main__":
    main()
