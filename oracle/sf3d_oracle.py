"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, fp32) of the Stable Fast 3D mesh
path.  Never imported by the product package; only tests/, smoke() and bench.py's CPU legs.

Pinned: yes -- tests/test_oracle_golden.py checks every function against
tests/golden/sf3d_*.npz, produced by the UNMODIFIED reference (oracle/make_golden_sf3d.py
runs sf3d.system.SF3D.query_triplane / triplane_to_meshes, sf3d.models.network.MaterialMLP and
sf3d.models.isosurface.MarchingTetrahedraHelper from /root/reference through oracle/ref_shim.py).

What each function restates (paths relative to /root/reference/StableFast):
  deform_grid         sf3d/models/isosurface.py:106-113, 210-213
  marching_tets       sf3d/models/isosurface.py:144-203 (sort_edges :135-142, tables :30-69)
  triplane_to_mesh    sf3d/system.py:141-168 (query :170-198 and heads via oracle/field_oracle.py)
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import field_oracle as fo

F32 = np.float32

TRIANGLE_TABLE = np.array(  # isosurface.py:30-55
    [[-1, -1, -1, -1, -1, -1], [1, 0, 2, -1, -1, -1], [4, 0, 3, -1, -1, -1], [1, 4, 2, 1, 3, 4],
     [3, 1, 5, -1, -1, -1], [2, 3, 0, 2, 5, 3], [1, 4, 0, 1, 5, 4], [4, 2, 5, -1, -1, -1],
     [4, 5, 2, -1, -1, -1], [4, 1, 0, 4, 5, 1], [3, 2, 0, 3, 5, 2], [1, 3, 5, -1, -1, -1],
     [4, 1, 2, 4, 3, 1], [3, 0, 4, -1, -1, -1], [2, 0, 1, -1, -1, -1], [-1, -1, -1, -1, -1, -1]], dtype=np.int64)
NUM_TRIANGLES = np.array([0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0], dtype=np.int64)  # :57-63
BASE_TET_EDGES = np.array([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3], dtype=np.int64)  # :67


def deform_grid(grid: np.ndarray, offsets: np.ndarray, resolution: int, points_range=(0, 1)) -> np.ndarray:
    scale = F32((points_range[1] - points_range[0]) / resolution)
    return (grid.astype(F32) + (scale * np.tanh(offsets.astype(F32))).astype(F32)).astype(F32)


def marching_tets(pos: np.ndarray, sdf: np.ndarray, tets: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(verts (V,3) f32, faces (F,3) i64) exactly as _forward orders them."""
    pos = pos.astype(F32)
    sdf = sdf.astype(F32).reshape(-1)
    occ = sdf > 0
    occ4 = occ[tets.reshape(-1)].reshape(-1, 4)
    occ_sum = occ4.sum(-1)
    valid = (occ_sum > 0) & (occ_sum < 4)
    edges = tets[valid][:, BASE_TET_EDGES].reshape(-1, 2)
    edges = np.sort(edges, axis=1)  # sort_edges: smaller index first
    uniq, inverse = np.unique(edges, axis=0, return_inverse=True)  # lexicographic, like torch.unique(dim=0)
    inverse = inverse.reshape(-1)
    mask = occ[uniq.reshape(-1)].reshape(-1, 2).sum(-1) == 1
    mapping = -np.ones(uniq.shape[0], dtype=np.int64)
    mapping[mask] = np.arange(mask.sum(), dtype=np.int64)
    idx_map = mapping[inverse].reshape(-1, 6)
    iv = uniq[mask]
    p = pos[iv.reshape(-1)].reshape(-1, 2, 3)
    s = sdf[iv.reshape(-1)].reshape(-1, 2, 1).copy()
    s[:, -1] *= F32(-1)
    den = s.sum(1, keepdims=True, dtype=F32)
    w = (s[:, ::-1] / den).astype(F32)
    verts = ((p * w).astype(F32)).sum(1, dtype=F32)
    code = (occ4[valid] * (2 ** np.arange(4))).sum(-1)
    ntri = NUM_TRIANGLES[code]
    one, two = ntri == 1, ntri == 2
    f1 = np.take_along_axis(idx_map[one], TRIANGLE_TABLE[code[one]][:, :3], axis=1).reshape(-1, 3)
    f2 = np.take_along_axis(idx_map[two], TRIANGLE_TABLE[code[two]][:, :6], axis=1).reshape(-1, 3)
    return verts.astype(F32), np.concatenate([f1, f2], axis=0).astype(np.int64)


def heads_from_state_dict(sd, name: str):
    ws = [np.asarray(sd[f"heads.{name}.{i}.weight"], dtype=F32) for i in (0, 2, 4)]
    bs = [np.asarray(sd[f"heads.{name}.{i}.bias"], dtype=F32) for i in (0, 2, 4)]
    return ws, bs


def triplane_to_mesh(
    triplane: np.ndarray, sd: Dict[str, np.ndarray], grid: np.ndarray, tets: np.ndarray, resolution: int,
    threshold: float, radius: float = 0.87, density_out_bias: float = -1.0,
) -> Dict[str, np.ndarray]:
    """One element of SF3D.triplane_to_meshes (system.py:141-168)."""
    bbox_lo, bbox_hi = F32(-radius), F32(radius)
    g = grid.astype(F32)
    positions = ((g - F32(0)) / F32(1 - 0)) * (bbox_hi - bbox_lo) + bbox_lo  # scale_tensor with the bbox tensor
    feats = fo.sf3d_query_triplane(positions, triplane, radius)
    density = fo.material_mlp_head(feats, *heads_from_state_dict(sd, "density"), out_bias=density_out_bias, activation="trunc_exp")
    offset = fo.material_mlp_head(feats, *heads_from_state_dict(sd, "vertex_offset"))
    sdf = (density - F32(threshold)).astype(F32)
    gdef = deform_grid(g, offset, resolution)
    v, f = marching_tets(gdef, sdf, tets)
    v = ((v - F32(0)) / F32(1 - 0)) * (bbox_hi - bbox_lo) + bbox_lo
    return dict(features=feats, density=density, vertex_offset=offset, sdf=sdf, grid_vertices=gdef, v_pos=v.astype(F32), t_pos_idx=f)
