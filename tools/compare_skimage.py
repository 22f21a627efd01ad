#!/usr/bin/env python3
"""Compare the in-repo marching cubes with the reference's REAL dependency, scikit-image's Lewiner
``skimage.measure.marching_cubes`` (/root/reference/TripoSR/tsr/models/isosurface.py:7,46-48), on a host where
scikit-image is installed (it is not in this repository's build environment: SURVEY 8c, DESIGN 2 -- parity unpinned).

    python tools/compare_skimage.py [--volume sphere|torus|gyroid|smooth|noise|field] [--R 64] [--npz grid.npz]

For each volume it runs  level -> skimage.measure.marching_cubes(level, 0.0)  exactly as MarchingCubeHelper.forward does
(then faces[:, [1,0,2]] and verts/(R-1)) and the in-repo CPU oracle (oracle/mc_oracle.c, the algorithm the CUDA kernels
are bit-exact against) and reports what can differ between a 33-case Lewiner table and the classic 256-case table:
  * vertex / face counts, Euler characteristic, connected components, signed volume of both meshes;
  * the vertex SETS (every vertex lies on a lattice edge: matched by (edge, position) within 2e-6 of a cell);
  * how many lattice cells hold a different number of triangles, split by whether the cell's cube case is
    face-ambiguous (the only cells where the two algorithms may legitimately choose different topology);
  * the symmetric Hausdorff distance between the two triangle soups (sampled at vertices).
Exit status 0 = identical geometry up to vertex numbering on non-ambiguous cells; the report says by how much the
ambiguous cells differ.  No GPU needed.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_volume(kind: str, R: int, seed: int = 0) -> np.ndarray:
    a = np.linspace(-1, 1, R, dtype=np.float32)
    x, y, z = np.meshgrid(a, a, a, indexing="ij")
    if kind == "sphere":
        return (0.6 - np.sqrt(x * x + y * y + z * z)).astype(np.float32)
    if kind == "torus":
        q = np.sqrt(x * x + y * y) - 0.55
        return (0.22 - np.sqrt(q * q + z * z)).astype(np.float32)
    if kind == "gyroid":
        s = 3.0 * np.pi
        return (np.sin(s * x) * np.cos(s * y) + np.sin(s * y) * np.cos(s * z) + np.sin(s * z) * np.cos(s * x)).astype(np.float32)
    rng = np.random.RandomState(seed)
    v = rng.randn(R, R, R).astype(np.float32)
    if kind == "smooth":
        for ax in range(3):
            v = (np.roll(v, 1, ax) + v + np.roll(v, -1, ax)) / 3
    return v.astype(np.float32)


def field_volume(R: int) -> np.ndarray:
    """density_act - median of the benchmark's baked field through the CPU oracle (slow beyond R ~ 96)."""
    import torch

    from bench import baked_triplane, decoder_numpy
    from oracle import field_oracle as fo

    _, ws, bs = decoder_numpy(0)
    d = fo.grid_density(R, baked_triplane(100).numpy(), ws, bs)
    return (d - np.median(d)).astype(np.float32)


def face_ambiguous_cases() -> np.ndarray:
    amb = np.zeros(256, bool)
    for case in range(256):
        s = [(case >> c) & 1 for c in range(8)]
        for axis in range(3):
            for side in (0, 1):
                cs = [c for c in range(8) if ((c >> (2 - axis)) & 1) == side]
                v = [s[c] for c in cs]  # ordered (0,0),(0,1),(1,0),(1,1) in the two other axes
                if v[0] == v[3] and v[1] == v[2] and v[0] != v[1]:
                    amb[case] = True
    return amb


def cube_cases(level: np.ndarray) -> np.ndarray:
    p = level > 0
    c = np.zeros(tuple(n - 1 for n in level.shape), np.uint8)
    for di in (0, 1):
        for dj in (0, 1):
            for dk in (0, 1):
                c |= p[di : di + c.shape[0], dj : dj + c.shape[1], dk : dk + c.shape[2]].astype(np.uint8) << (4 * di + 2 * dj + dk)
    return c


def mesh_stats(v: np.ndarray, f: np.ndarray):
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    ne = len(np.unique(e, axis=0))
    chi = len(v) - ne + len(f)
    a, b, c = v[f[:, 0]].astype(np.float64), v[f[:, 1]].astype(np.float64), v[f[:, 2]].astype(np.float64)
    vol = float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0)
    return dict(verts=len(v), faces=len(f), euler=int(chi), signed_volume=vol)


def tris_per_cell(v_idx: np.ndarray, f: np.ndarray, shape) -> np.ndarray:
    """Triangles per lattice cell (v_idx in index units): the cell of a triangle = floor of its centroid."""
    cen = v_idx[f].mean(axis=1)
    ijk = np.clip(np.floor(cen + 1e-7).astype(int), 0, np.array(shape) - 2)
    out = np.zeros(tuple(n - 1 for n in shape), np.int32)
    np.add.at(out, (ijk[:, 0], ijk[:, 1], ijk[:, 2]), 1)
    return out


def vertex_keys(v_idx: np.ndarray) -> np.ndarray:
    return np.unique(np.round(v_idx.astype(np.float64) * 2**18).astype(np.int64), axis=0)


def hausdorff(a: np.ndarray, b: np.ndarray) -> float:
    from scipy.spatial import cKDTree

    return float(max(cKDTree(b).query(a)[0].max(), cKDTree(a).query(b)[0].max()))


def compare(level: np.ndarray, name: str) -> bool:
    from skimage import measure

    from oracle import mc_oracle

    R = level.shape[0]
    v_s, f_s, _, _ = measure.marching_cubes(np.ascontiguousarray(level), 0.0)  # isosurface.py:46-48
    f_s = f_s[:, [1, 0, 2]].astype(np.int64)  # :52
    v_o, f_o, _ = mc_oracle.marching_cubes_slab(level, sub=np.float32(0.0), flags=1)  # FLIP only: index units like skimage's verts
    amb = face_ambiguous_cases()[cube_cases(level)]
    ts, to = tris_per_cell(v_s, f_s, level.shape), tris_per_cell(v_o, f_o, level.shape)
    diff = ts != to
    ks, ko = vertex_keys(v_s), vertex_keys(v_o)
    common = len(set(map(tuple, ks)) & set(map(tuple, ko)))
    print(f"== {name} R={R}")
    print("   skimage (Lewiner):", mesh_stats(v_s / (R - 1), f_s))
    print("   in-repo (classic):", mesh_stats(v_o / (R - 1), f_o))
    print(f"   vertex positions: {len(ks)} vs {len(ko)}, {common} coincide within 4e-6 cell")
    print(f"   cells with a different triangle count: {int(diff.sum())} of {int((ts > 0).sum())} active "
          f"({int((diff & amb).sum())} in face-ambiguous cells, {int((diff & ~amb).sum())} elsewhere); face-ambiguous active cells: {int((amb & (to > 0)).sum())}")
    print(f"   symmetric Hausdorff distance (vertices): {hausdorff(v_s, v_o):.3e} cells")
    return int((diff & ~amb).sum()) == 0 and common == len(ko) == len(ks)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--volume", default="all")
    ap.add_argument("--R", type=int, default=64)
    ap.add_argument("--npz", default=None, help="npz with a 3-D float32 array 'level' (e.g. density_act - threshold dumped from a GPU run)")
    args = ap.parse_args()
    try:
        import skimage  # noqa: F401
    except ImportError:
        print("scikit-image is not installed here: run this tool where the reference's dependency exists (pip install scikit-image)")
        return 2
    ok = True
    if args.npz:
        ok &= compare(np.load(args.npz)["level"].astype(np.float32), args.npz)
    else:
        kinds = ["sphere", "torus", "gyroid", "smooth", "noise", "field"] if args.volume == "all" else [args.volume]
        for k in kinds:
            lvl = field_volume(min(args.R, 64)) if k == "field" else make_volume(k, args.R)
            ok &= compare(lvl, k)
    print("RESULT:", "identical outside face-ambiguous cells" if ok else "differences outside face-ambiguous cells (see above)")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
