"""Drop-in for /root/reference/StableFast/sf3d/models/isosurface.py (MarchingTetrahedraHelper).

``forward(level, deformation) -> Mesh`` keeps the reference contract -- level (Nv,1),
deformation (Nv,3) or None, float vertices and long faces on ``level.device``, the same
vertex numbering (lexicographic order of the sorted unique crossing edges, :153-168) and the
same face order (all 1-triangle tets, then all 2-triangle tets, :187-201) -- but runs as a
flag + prefix scan over the grid's static, pre-sorted edge list (csrc/sf3d.cu) instead of a
per-call ``torch.unique`` sort.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from ... import runtime
from .mesh import Mesh


def detect_lattice(vertices: np.ndarray):
    """Is ``vertices`` (N,3) a lattice in index order -- vertices[(a*nB + b)*nC + c] = (coordinates picked from three
    lists by a, b, c), for some assignment of the spatial axes to the slow / mid / fast index?  Returns
    ``(extents, spatial_dim, coords)`` (coords[k]: float32 array of index k's coordinate values) or None.  The check is
    exact (bit-for-bit) and costs one pass over the array; a Kuhn / marching-cubes-style tet grid passes, a quartet
    (BCC) tetrahedralisation does not and keeps the arbitrary-position kernel."""
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    n = v.shape[0]
    if v.ndim != 2 or v.shape[1] != 3 or n < 8:
        return None
    d1 = np.flatnonzero(v[1] != v[0])
    if d1.size != 1:
        return None
    fast = int(d1[0])
    same = np.flatnonzero(v[:, fast] == v[0, fast])  # the fast coordinate returns to its first value every nC vertices
    if same.size < 2:
        return None
    nC = int(same[1])
    if nC < 2 or n % nC or nC >= n:
        return None
    d2 = np.flatnonzero(v[nC] != v[0])
    if d2.size != 1 or int(d2[0]) == fast:
        return None
    mid = int(d2[0])
    slow = 3 - fast - mid
    rows = v[::nC]
    same = np.flatnonzero(rows[:, mid] == rows[0, mid])
    if same.size < 2:
        return None
    nB = int(same[1])
    if nB < 2 or (n // nC) % nB:
        return None
    nA = n // (nC * nB)
    if nA < 2:
        return None
    g = v.reshape(nA, nB, nC, 3)
    ca, cb, cc = g[:, 0, 0, slow].copy(), g[0, :, 0, mid].copy(), g[0, 0, :, fast].copy()
    ok = (
        np.array_equal(g[..., slow], np.broadcast_to(ca[:, None, None], (nA, nB, nC)))
        and np.array_equal(g[..., mid], np.broadcast_to(cb[None, :, None], (nA, nB, nC)))
        and np.array_equal(g[..., fast], np.broadcast_to(cc[None, None, :], (nA, nB, nC)))
    )
    if not ok:
        return None
    return (nA, nB, nC), (slow, mid, fast), (ca, cb, cc)


class IsosurfaceHelper(nn.Module):
    points_range: Tuple[float, float] = (0, 1)

    @property
    def grid_vertices(self) -> torch.Tensor:
        raise NotImplementedError

    @property
    def requires_instance_per_batch(self) -> bool:
        return False


class MarchingTetrahedraHelper(IsosurfaceHelper):
    def __init__(self, resolution: int, tets_path: str):
        super().__init__()
        self.resolution = resolution
        self.tets_path = tets_path
        tets = np.load(self.tets_path)
        self.register_buffer("_grid_vertices", torch.from_numpy(tets["vertices"]).float(), persistent=False)
        self.register_buffer("indices", torch.from_numpy(tets["indices"]).long(), persistent=False)
        self._all_edges: Optional[torch.Tensor] = None
        self._topology = None  # (device, edges int32 (E,2), tets int32 (T,4), tet_edges int32 (T,6))
        # (extents of the slow / mid / fast vertex index, the spatial axis of each, their coordinate lists) when the vertex
        # array is an outer product of three coordinate lists, else None
        self.lattice = detect_lattice(tets["vertices"])

    def normalize_grid_deformation(self, grid_vertex_offsets: torch.Tensor) -> torch.Tensor:
        """isosurface.py:106-113 (eager form, kept for callers outside the fused path)."""
        return (self.points_range[1] - self.points_range[0]) / self.resolution * torch.tanh(grid_vertex_offsets)

    @property
    def grid_vertices(self) -> torch.Tensor:
        return self._grid_vertices

    @property
    def all_edges(self) -> torch.Tensor:
        """Sorted unique edges of the grid (isosurface.py:119-133), int64 (E,2)."""
        if self._all_edges is None:
            base = torch.tensor([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3], dtype=torch.long, device=self.indices.device)
            e = self.indices[:, base].reshape(-1, 2)
            e = torch.sort(e, dim=1)[0]
            nv = self._grid_vertices.shape[0]
            keys = torch.unique(e[:, 0] * nv + e[:, 1])  # lexicographic order of (a,b) == order of a*Nv+b
            self._all_edges = torch.stack([keys // nv, keys % nv], dim=-1)
        return self._all_edges

    def topology(self, device: torch.device):
        """Static index arrays of the grid on ``device`` (built once, cached)."""
        if self._topology is None or self._topology[0] != device:
            if self.indices.device != device:
                self.indices = self.indices.to(device)
                self._all_edges = None
            edges = self.all_edges
            nv = self._grid_vertices.shape[0]
            if nv >= 2**31 or edges.shape[0] >= 2**31:
                raise NotImplementedError("tet grids with >= 2^31 vertices or edges are not supported")
            base = torch.tensor([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3], dtype=torch.long, device=device)
            te = torch.sort(self.indices[:, base].reshape(-1, 2), dim=1)[0]
            tet_edges = torch.searchsorted(edges[:, 0] * nv + edges[:, 1], te[:, 0] * nv + te[:, 1]).view(-1, 6)
            self._topology = (
                device, edges.to(torch.int32).contiguous(), self.indices.to(torch.int32).contiguous(),
                tet_edges.to(torch.int32).contiguous(),
            )
        return self._topology[1:]

    def forward(self, level: torch.Tensor, deformation: Optional[torch.Tensor] = None, v_pos_affine=None) -> Mesh:
        """``v_pos_affine`` (optional, 8 floats, see ``runtime.marching_tets``): the caller's scale_tensor of ``v_pos``
        folded into the vertex kernel (sf3d/system.py:162-164)."""
        dev = level.device
        base = self._grid_vertices.to(dev)
        if deformation is not None:
            scale = (self.points_range[1] - self.points_range[0]) / self.resolution
            grid_vertices = runtime.mtet_deform(base, deformation.detach().to(torch.float32).view(-1, 3), scale)
        else:
            grid_vertices = base
        edges, tets, tet_edges = self.topology(dev)
        sdf = level.detach().to(torch.float32).contiguous().view(-1)
        v_pos, t_pos_idx = runtime.marching_tets(grid_vertices, sdf, edges, tets, tet_edges, affine=v_pos_affine)
        return Mesh(
            v_pos=v_pos, t_pos_idx=t_pos_idx,
            grid_vertices=grid_vertices, tet_edges=self.all_edges, grid_level=level, grid_deformation=deformation,
        )
