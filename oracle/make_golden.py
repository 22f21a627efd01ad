#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the UNMODIFIED
reference (imported from /root/reference through oracle/ref_shim.py).

The reference cannot travel to the GPU box, so its outputs do.  Re-run with
    python oracle/make_golden.py
whenever the fixtures need regenerating (only possible where /root/reference exists).

Fixtures
  lattice.npz        MarchingCubeHelper.grid_vertices + the two scale_tensor remaps
                     (isosurface.py:25-39, system.py:177-181, nerf_renderer.py:52-54)
  field_small.npz    query_triplane on a 3x40x16x16 triplane, lattice + random +
                     border/out-of-range positions (nerf_renderer.py:41-91)
  field_64.npz       same on a seeded 3x40x64x64 triplane (inputs regenerated from the
                     seed, checksummed), random positions
  helper_sphere.npz  MarchingCubeHelper.forward wrapper semantics with
                     skimage.measure.marching_cubes replaced by the in-repo oracle MC
                     (skimage is not installed; isosurface.py:41-54)
  extract_mesh.npz   the reference's own TSR.extract_mesh body (system.py:171-200) run
                     on decoder+renderer+helper, oracle MC patched in, sink captured
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mc_oracle, ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
RADIUS = 0.87


def baked_triplane(seed: int, H: int, W: int, noise: float = 0.05) -> torch.Tensor:
    """SURVEY 8d family B: plane[p,c,h,w] = A[p,c]*(u_w^2+v_h^2)/2 + noise*randn."""
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(3, 40, 1, 1, generator=g)
    u = ((torch.arange(W) + 0.5) / W * 2 - 1).view(1, 1, 1, W)
    v = ((torch.arange(H) + 0.5) / H * 2 - 1).view(1, 1, H, 1)
    return (A * (u * u + v * v) / 2 + noise * torch.randn(3, 40, H, W, generator=g)).float()


def sd_arrays(decoder):
    sd = decoder.state_dict()
    out = {}
    for i in range(0, 20, 2):
        out[f"w{i // 2}"] = sd[f"layers.{i}.weight"].numpy().copy()
        out[f"b{i // 2}"] = sd[f"layers.{i}.bias"].numpy().copy()
    return out


def main() -> None:
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_shim.load_triposr()
    scale_tensor = ref.utils.scale_tensor

    # ---------------------------------------------------------------- lattice
    lat = {}
    for R in (2, 5, 16, 33, 64):
        h = ref.isosurface.MarchingCubeHelper(R)
        gv = h.grid_vertices
        lat[f"axis_{R}"] = torch.linspace(0, 1, R).numpy()
        a1 = scale_tensor(torch.linspace(0, 1, R), (0, 1), (-RADIUS, RADIUS))
        a2 = scale_tensor(a1, (-RADIUS, RADIUS), (-1, 1))
        lat[f"axis_scaled_{R}"] = a1.numpy()
        lat[f"axis_unit_{R}"] = a2.numpy()
        if R <= 16:
            lat[f"verts_{R}"] = gv.numpy()
            lat[f"verts_unit_{R}"] = scale_tensor(scale_tensor(gv, (0, 1), (-RADIUS, RADIUS)), (-RADIUS, RADIUS), (-1, 1)).numpy()
    np.savez_compressed(os.path.join(GOLD, "lattice.npz"), **lat)

    # ------------------------------------------------------------ field_small
    dec = ref_shim.make_reference_decoder(0)
    rend = ref_shim.make_reference_renderer(8192)
    tp = baked_triplane(1, 16, 16, noise=0.3)
    R = 12
    h = ref.isosurface.MarchingCubeHelper(R)
    pos_lat = scale_tensor(h.grid_vertices, (0, 1), (-RADIUS, RADIUS))
    g = torch.Generator().manual_seed(7)
    pos_rand = (torch.rand(1500, 3, generator=g) * 2 - 1) * RADIUS
    edge = torch.tensor(
        [[-RADIUS, -RADIUS, -RADIUS], [RADIUS, RADIUS, RADIUS], [RADIUS, -RADIUS, 0.0], [0.0, 0.0, 0.0],
         [-1.0, 0.3, 0.9], [1.2, -1.3, 0.1], [0.869999, -0.869999, 0.5], [2.0, 2.0, 2.0]], dtype=torch.float32)
    pos = torch.cat([pos_lat, pos_rand, edge], 0)
    with torch.no_grad():
        out = rend.query_triplane(dec, pos, tp)
    np.savez_compressed(
        os.path.join(GOLD, "field_small.npz"), triplane=tp.numpy(), positions=pos.numpy(), n_lattice=np.int64(R),
        **{k: v.numpy() for k, v in out.items()}, **sd_arrays(dec),
    )

    # --------------------------------------------------------------- field_64
    dec1 = ref_shim.make_reference_decoder(1)
    torch.manual_seed(11)
    tp64 = torch.randn(3, 40, 64, 64)
    g = torch.Generator().manual_seed(12)
    pos64 = (torch.rand(4000, 3, generator=g) * 2 - 1) * RADIUS
    with torch.no_grad():
        out64 = rend.query_triplane(dec1, pos64, tp64)
    np.savez_compressed(
        os.path.join(GOLD, "field_64.npz"), triplane_seed=np.int64(11),
        triplane_sha256=np.frombuffer(hashlib.sha256(tp64.numpy().tobytes()).digest(), dtype=np.uint8),
        positions=pos64.numpy(), **{k: v.numpy() for k, v in out64.items()}, **sd_arrays(dec1),
    )

    # ---------------------------------------------------------- helper_sphere
    import skimage.measure as skm  # the shim's stub module

    skm.marching_cubes = mc_oracle.marching_cubes  # in-repo oracle MC stands in for scikit-image
    R = 16
    h = ref.isosurface.MarchingCubeHelper(R)
    gv = h.grid_vertices * 2 - 1
    density = (0.6 - gv.norm(dim=-1)).view(-1, 1)  # > 0 inside
    v_pos, t_idx = h(-density)  # caller passes -(density - thr) (system.py:184)
    np.savez_compressed(
        os.path.join(GOLD, "helper_sphere.npz"), level_in=(-density).numpy(), v_pos=v_pos.numpy(), t_pos_idx=t_idx.numpy(),
        v_dtype=str(v_pos.dtype), t_dtype=str(t_idx.dtype), resolution=np.int64(R),
    )

    # ----------------------------------------------------------- extract_mesh
    import tsr.system as ref_system  # the reference's own module (bpy stubbed)

    captured = []

    class _Host(torch.nn.Module):
        """Bare carrier for the three sub-modules extract_mesh touches; the method body
        that runs is the reference's (system.py:118-124,171-200)."""

        set_marching_cubes_resolution = ref_system.TSR.set_marching_cubes_resolution
        extract_mesh = ref_system.TSR.extract_mesh

        def __init__(self, decoder, renderer):
            super().__init__()
            self.decoder, self.renderer, self.isosurface_helper = decoder, renderer, None

        def import_obj_blender(self, verts, faces, vertex_colors=None, name="NewMesh"):
            captured.append((verts, faces, vertex_colors, name))

    host = _Host(dec, rend)
    tpB = baked_triplane(3, 16, 16, noise=0.05)
    R = 24
    with torch.no_grad():
        dens = rend.query_triplane(
            dec, scale_tensor(ref.isosurface.MarchingCubeHelper(R).grid_vertices, (0, 1), (-RADIUS, RADIUS)), tpB
        )["density_act"]
    thr = float(dens.median())
    host.extract_mesh(tpB[None], enable_texture=True, mesh_name="golden", resolution=R, threshold=thr)
    verts, faces, colors, name = captured[0]
    np.savez_compressed(
        os.path.join(GOLD, "extract_mesh.npz"), triplane=tpB.numpy(), resolution=np.int64(R), threshold=np.float64(thr),
        density_act=dens.numpy().reshape(R, R, R), verts=verts, faces=faces, colors=colors, **sd_arrays(dec),
    )
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))
    print("extract_mesh golden:", verts.shape, faces.shape, verts.dtype, faces.dtype, "thr", thr)


if __name__ == "__main__":
    main()
