"""TEST INFRASTRUCTURE ONLY -- numpy (fp32) restatement of the texture baker's rasterise / interpolate step.
Never imported by the product package; only tests/ use it, as the checker.

Pinned: yes, to the reference's own Python functions of the same name -- ``rasterize_cpu`` / ``interpolate_cpu``,
/root/reference/StableFast/sf3d/texture_baker/common.py:104-142,214-230, the routines the production path calls inside
texture_baker.dll (baker.py:30-57,92-118; Windows-only binary, no source) -- through ``tests/golden/bake.npz``
(``oracle/make_golden_bake.py``).  Difference by construction: the reference finds the covering triangle with a BVH and
returns the FIRST hit of its traversal; a texel covered by several triangles (shared edges / vertices of the atlas) goes
to the LOWEST triangle index here and in the CUDA kernel.  Barycentrics and interpolated attributes of every texel whose
triangle agrees are bit-identical.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def texel_points(resolution: int):
    """common.py:133: texel (y, x) -> point (x / width, 1.0 - y / height), Python floats rounded to fp32 on first use."""
    k = np.arange(resolution, dtype=np.float64)
    return (k / resolution).astype(F32), (1.0 - k / resolution).astype(F32)  # px[x], py[y]


def barycentric(px, py, v0, v1, v2):
    """common.py:104-121 on fp32 scalars (every operation rounded to fp32)."""
    ax, ay = v1[0] - v0[0], v1[1] - v0[1]
    bx, by = v2[0] - v0[0], v2[1] - v0[1]
    qx, qy = px - v0[0], py - v0[1]
    d00 = ax * ax + ay * ay
    d01 = ax * bx + ay * by
    d11 = bx * bx + by * by
    d20 = qx * ax + qy * ay
    d21 = qx * bx + qy * by
    denom = d00 * d11 - d01 * d01
    with np.errstate(all="ignore"):
        v = (d11 * d20 - d01 * d21) / denom
        w = (d00 * d21 - d01 * d20) / denom
    u = F32(1.0) - v - w
    return u, v, w


def rasterize(uv: np.ndarray, faces: np.ndarray, resolution: int, window=None) -> np.ndarray:
    """(res, res, 4) fp32: (u, v, w, triangle index) or (0, 0, 0, -1); lowest covering triangle index wins.
    ``window=(y0, y1, x0, x1)`` restricts the evaluation to texels [y0:y1, x0:x1] (returns that crop)."""
    uv = uv.astype(F32)
    px, py = texel_points(resolution)
    if window is not None:
        py, px = py[window[0] : window[1]], px[window[2] : window[3]]
    PX, PY = np.meshgrid(px, py, indexing="xy")  # [y, x]
    out = np.zeros((len(py), len(px), 4), F32)
    out[..., 3] = -1
    for f in range(len(faces) - 1, -1, -1):  # descending: lower indices overwrite
        v0, v1, v2 = uv[faces[f, 0]], uv[faces[f, 1]], uv[faces[f, 2]]
        u, v, w = barycentric(PX, PY, v0, v1, v2)
        inside = (u >= 0) & (v >= 0) & (w >= 0)
        out[inside] = np.stack([u[inside], v[inside], w[inside], np.full(int(inside.sum()), f, F32)], -1)
    return out


def interpolate(attr: np.ndarray, faces: np.ndarray, rast: np.ndarray) -> np.ndarray:
    """common.py:214-230: attr[i0] * u + attr[i1] * v + attr[i2] * w in fp32, zero where no triangle."""
    attr = attr.astype(F32)
    tri = rast[..., 3].astype(np.int64)
    ok = rast[..., 3] >= 0
    t = np.where(ok, tri, 0)
    i0, i1, i2 = faces[t, 0], faces[t, 1], faces[t, 2]
    val = attr[i0] * rast[..., 0:1] + attr[i1] * rast[..., 1:2] + attr[i2] * rast[..., 2:3]
    return np.where(ok[..., None], val, F32(0)).astype(F32)
