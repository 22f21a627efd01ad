#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- generates tests/golden/render_rays.npz by running the UNMODIFIED reference:
TriplaneNeRFRenderer.forward / _forward (/root/reference/TripoSR/tsr/models/nerf_renderer.py:93-172) and
rays_intersect_bbox (tsr/utils.py:115-149) on a 16x16 baked triplane, 128 samples per ray (config.yaml:33),
rays that all hit the +-0.87 box (the reference's _forward raises a shape error as soon as one ray misses:
z_vals is built from the valid rays only, nerf_renderer.py:106-117, and then added to ALL ray origins),
including axis-aligned directions with zero components (rays_d_valid, utils.py:124-126).  The decoder's
density row of the last layer is scaled and biased so that opacities span (0, 1): with the random-init values
every alpha would be ~1e-4 and the composite would only test the background term.

    python oracle/make_golden_render.py        (needs /root/reference; run in the build container)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.make_golden import baked_triplane, sd_arrays  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def make_rays(n: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * (1.6 + 0.8 * torch.rand(n, 1, generator=g))
    target = (torch.rand(n, 3, generator=g) * 2 - 1) * 0.8  # inside the box: every ray hits it
    d = target - o
    d = d / d.norm(dim=-1, keepdim=True)
    # axis-aligned rays: two direction components exactly zero (rays_d_valid replaces them by 1e-6)
    o[:6] = torch.tensor([[2.0, 0.1, -0.2], [-2.0, 0.5, 0.5], [0.3, 2.0, 0.0], [0.0, -2.0, 0.86], [0.2, 0.2, 2.0], [0.8, 0.0, -2.0]])
    d[:6] = torch.tensor([[-1.0, 0, 0], [1.0, 0, 0], [0, -1.0, 0], [0, 1.0, 0], [0, 0, -1.0], [0, 0, 1.0]])
    return o.float().contiguous(), d.float().contiguous()


def main() -> None:
    ref = ref_shim.load_triposr()
    dec = ref_shim.make_reference_decoder(3)
    with torch.no_grad():
        dec.layers[18].weight[0] *= 40.0  # spatial contrast of the density
        dec.layers[18].bias[0] -= 3.0
    rend = ref_shim.make_reference_renderer(8192)
    rend.eval()
    tp = baked_triplane(5, 16, 16, noise=0.3)
    rays_o, rays_d = make_rays(300, 21)
    with torch.no_grad():
        comp = rend(dec, tp, rays_o.view(15, 20, 3), rays_d.view(15, 20, 3))
        t_near, t_far, valid = ref.utils.rays_intersect_bbox(rays_o, rays_d, rend.cfg.radius)
        # batched form (triplane.ndim == 5): one _forward per scene
        tp2 = torch.stack([tp, tp.flip(-1)])
        comp2 = rend(dec, tp2, torch.stack([rays_o[:50], rays_o[50:100]]), torch.stack([rays_d[:50], rays_d[50:100]]))
    assert comp.shape == (15, 20, 3) and bool(valid.all())
    # one ray that misses: the reference raises (shape mismatch between the valid-ray z_vals and all ray origins)
    o_bad, d_bad = rays_o.clone(), rays_d.clone()
    d_bad[7] = o_bad[7] / o_bad[7].norm()
    try:
        with torch.no_grad():
            rend(dec, tp, o_bad, d_bad)
        raised = ""
    except RuntimeError as e:
        raised = type(e).__name__
    assert raised == "RuntimeError"
    op = 1.0 - comp.view(-1, 3)[valid.view(-1)].min(dim=-1).values
    print(f"rays {rays_o.shape[0]}, valid {int(valid.sum())}, composite range [{float(comp.min()):.3f}, {float(comp.max()):.3f}], "
          f"non-background rays {(op > 0.05).sum().item()}")
    np.savez_compressed(
        os.path.join(GOLD, "render_rays.npz"), triplane=tp.numpy(), rays_o=rays_o.numpy(), rays_d=rays_d.numpy(),
        comp_rgb=comp.numpy(), t_near=t_near.numpy(), t_far=t_far.numpy(), rays_valid=valid.numpy(), comp_rgb_batched=comp2.numpy(),
        num_samples=np.int64(rend.cfg.num_samples_per_ray), radius=np.float32(rend.cfg.radius), **sd_arrays(dec),
    )


if __name__ == "__main__":
    main()
