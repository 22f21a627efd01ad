#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/sf3d_*.npz by running the UNMODIFIED
reference's Stable Fast 3D mesh path (imported from /root/reference through oracle/ref_shim.py):
  sf3d.system.SF3D.query_triplane / triplane_to_meshes      StableFast/sf3d/system.py:141-198
  sf3d.models.network.MaterialMLP                           StableFast/sf3d/models/network.py:148-208
  sf3d.models.isosurface.MarchingTetrahedraHelper           StableFast/sf3d/models/isosurface.py:24-229
The reference's tet grid blob (load/tets/160_tets.npz) is missing from the checkout, so a
Kuhn grid from sculptmate_b200.sf3d.tets is used (sizes stored in the fixtures).

    python oracle/make_golden_sf3d.py          (only where /root/reference exists)
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from sculptmate_b200.sf3d.tets import save_tet_grid  # noqa: E402  (pure numpy helper, no CUDA)

GOLD = os.path.join(ROOT, "tests", "golden")
RADIUS = 0.87
HEADS = [
    dict(name="density", out_channels=1, out_bias=-1.0, n_hidden_layers=2, output_activation="trunc_exp"),
    dict(name="vertex_offset", out_channels=3, n_hidden_layers=2),
]


def baked_triplane(seed: int, H: int, W: int, noise: float) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(3, 40, 1, 1, generator=g)
    u = (torch.arange(W) / (W - 1) * 2 - 1).view(1, 1, 1, W)
    v = (torch.arange(H) / (H - 1) * 2 - 1).view(1, 1, H, 1)
    return (A * (u * u + v * v) / 2 + noise * torch.randn(3, 40, H, W, generator=g)).float()


def main() -> None:
    ref = ref_shim.load_sf3d_system()
    tmp = tempfile.mkdtemp()

    # ------------------------------------------------ marching tets on analytic fields
    n = 12
    path = save_tet_grid(os.path.join(tmp, "tets12.npz"), n)
    helper = ref.isosurface.MarchingTetrahedraHelper(n, path)
    gv = helper.grid_vertices
    out = {"n": np.int64(n)}
    g = torch.Generator().manual_seed(5)
    fields = {
        "sphere": 0.37 - (gv - 0.5).norm(dim=-1),
        "torus": 0.12 - ((((gv[:, :2] - 0.5).norm(dim=-1) - 0.28) ** 2 + (gv[:, 2] - 0.5) ** 2).sqrt()),
        "noise": torch.randn(gv.shape[0], generator=g) * 0.5,
    }
    for name, sdf in fields.items():
        deform = None if name != "torus" else torch.randn(gv.shape[0], 3, generator=g)
        with torch.no_grad():
            mesh = helper(sdf.view(-1, 1).float(), deform)
        out[f"{name}_sdf"] = sdf.float().numpy()
        if deform is not None:
            out[f"{name}_deform"] = deform.numpy()
            out[f"{name}_grid"] = mesh.extras["grid_vertices"].numpy()
        out[f"{name}_v"] = mesh.v_pos.numpy()
        out[f"{name}_f"] = mesh.t_pos_idx.numpy()
        print(name, mesh.v_pos.shape, mesh.t_pos_idx.shape)
    out["all_edges"] = helper.all_edges.numpy()
    np.savez_compressed(os.path.join(GOLD, "sf3d_mtet.npz"), **out)

    # ---------------------------------------- query_triplane + MaterialMLP + triplane_to_meshes
    torch.manual_seed(3)
    dec = ref.network.MaterialMLP(dict(in_channels=120, n_neurons=64, activation="silu", heads=[ref.network.HeadSpec(**h) for h in HEADS]))
    n = 10
    path = save_tet_grid(os.path.join(tmp, "tets10.npz"), n)
    helper = ref.isosurface.MarchingTetrahedraHelper(n, path)
    tp = baked_triplane(9, 24, 24, noise=0.2)
    host = types.SimpleNamespace(
        cfg=types.SimpleNamespace(radius=RADIUS, isosurface_threshold=10.0), decoder=dec, isosurface_helper=helper,
        bbox=torch.as_tensor([[-RADIUS] * 3, [RADIUS] * 3], dtype=torch.float32),
    )
    SF3D = ref.system.SF3D
    host.query_triplane = types.MethodType(SF3D.query_triplane, host)
    g = torch.Generator().manual_seed(4)
    pos = torch.cat([
        (torch.rand(600, 3, generator=g) * 2 - 1) * RADIUS,
        torch.tensor([[-RADIUS] * 3, [RADIUS] * 3, [0.0, 0.0, 0.0], [RADIUS, -RADIUS, 0.25], [1.1, -1.3, 0.2]], dtype=torch.float32),
    ])
    with torch.no_grad():
        feats = host.query_triplane(pos, tp)  # (1,N,120)
        decoded = dec(feats, include=["vertex_offset", "density"])
        # threshold inside the density range of a random-init decoder
        grid_pos = ref.system.scale_tensor(helper.grid_vertices, helper.points_range, host.bbox)
        dens_grid = dec(host.query_triplane(grid_pos, tp), include=["density"])["density"]
        host.cfg.isosurface_threshold = float(dens_grid.median())
        meshes = SF3D.triplane_to_meshes(host, tp[None])
    m = meshes[0]
    print("triplane_to_meshes:", m.v_pos.shape, m.t_pos_idx.shape, "thr", host.cfg.isosurface_threshold)
    sd = {k: v.numpy() for k, v in dec.state_dict().items()}
    np.savez_compressed(
        os.path.join(GOLD, "sf3d_path.npz"), n=np.int64(n), triplane=tp.numpy(), positions=pos.numpy(),
        features=feats.numpy(), density=decoded["density"].numpy(), vertex_offset=decoded["vertex_offset"].numpy(),
        threshold=np.float64(host.cfg.isosurface_threshold), v_pos=m.v_pos.numpy(), t_pos_idx=m.t_pos_idx.numpy(),
        grid_vertices=m.extras["grid_vertices"].numpy(), grid_level=m.extras["grid_level"].numpy(),
        **{"sd." + k: v for k, v in sd.items()},
    )
    for f in sorted(os.listdir(GOLD)):
        if f.startswith("sf3d"):
            print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
