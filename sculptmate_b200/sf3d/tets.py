"""Synthetic tetrahedral grids.

The reference loads ``load/tets/160_tets.npz`` (sf3d/models/isosurface.py:71-81,
sf3d/system.py:125-134); that blob is not part of the reference checkout
(.MISSING_LARGE_BLOBS), so its exact vertex/tet layout is unknown here.  The path itself
takes ANY npz with ``vertices`` (Nv,3) float in [0,1] and ``indices`` (Nt,4) int, so tests
and benchmarks use a Kuhn/Freudenthal grid: an n^3 lattice of cubes, 6 tetrahedra per
cube around the main diagonal.  Sizes are stated wherever a number is reported.
"""
from __future__ import annotations

import itertools

import numpy as np


def kuhn_tet_grid(n: int):
    """(vertices (n+1)^3 x 3 float32 in [0,1], indices 6n^3 x 4 int64)."""
    ax = np.linspace(0.0, 1.0, n + 1, dtype=np.float32)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    verts = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=-1).astype(np.float32)
    m = n + 1
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    base = (i * m + j) * m + k
    step = np.array([m * m, m, 1], dtype=np.int64)
    tets = []
    for perm in itertools.permutations(range(3)):  # one tet per monotone path 000 -> 111
        v0 = base
        v1 = v0 + step[perm[0]]
        v2 = v1 + step[perm[1]]
        v3 = v2 + step[perm[2]]
        # the marching-tets triangle table assumes one orientation for every tet: the path
        # volume det[e_p0, e_p1, e_p2] is the sign of the permutation, so odd ones are flipped
        odd = sum(perm[a] > perm[b] for a in range(3) for b in range(a + 1, 3)) % 2 == 1
        order = [v0, v2, v1, v3] if odd else [v0, v1, v2, v3]
        tets.append(np.stack(order, axis=-1).reshape(-1, 4))
    idx = np.stack(tets, axis=1).reshape(-1, 4).astype(np.int64)  # the 6 tets of a cube are adjacent
    return verts, idx


def save_tet_grid(path: str, n: int) -> str:
    v, t = kuhn_tet_grid(n)
    np.savez(path, vertices=v, indices=t)
    return path
