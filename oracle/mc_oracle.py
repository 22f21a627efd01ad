"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper around oracle/mc_oracle.c.

See mc_oracle.c for the algorithm, the canonical ordering and the statement that
MC parity is UNPINNED against scikit-image (absent here).  The wrapper semantics
follow /root/reference/TripoSR/tsr/models/isosurface.py:41-54.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsmb_oracle.so")

FLIP, DIV, AFFINE = 1, 2, 4


class Counts(ctypes.Structure):
    _fields_ = [
        ("nverts", ctypes.c_int64),
        ("ntris", ctypes.c_int64),
        ("npos", ctypes.c_int64),
        ("nneg", ctypes.c_int64),
        ("nverts_numbered", ctypes.c_int64),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mc_oracle.c")
    tab = os.path.join(_HERE, "mc_tables_oracle.h")
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(tab))
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/libsmb_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        f32p = ctypes.POINTER(ctypes.c_float)
        lib.smb_oracle_mc.restype = ctypes.c_int
        lib.smb_oracle_mc.argtypes = [
            f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float,
            f32p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(Counts),
        ]
        lib.smb_oracle_mc_cases.restype = None
        lib.smb_oracle_mc_cases.argtypes = [
            f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
            ctypes.POINTER(ctypes.c_ubyte),
        ]
        _lib = lib
    return _lib


def marching_cubes_slab(
    grid: np.ndarray,
    sub: float = 0.0,
    sign: float = 1.0,
    x_origin: int = 0,
    emit_last_plane: bool = True,
    flags: int = 0,
    vdiv: float = 1.0,
    vmul: float = 1.0,
    vadd: float = 0.0,
) -> Tuple[np.ndarray, np.ndarray, Counts]:
    """MC of val = (grid - sub) * sign at iso 0 over a (nx,ny,nz) fp32 slab."""
    lib = _load()
    g = np.ascontiguousarray(grid, dtype=np.float32)
    assert g.ndim == 3
    nx, ny, nz = g.shape
    f32p = ctypes.POINTER(ctypes.c_float)
    gp = g.ctypes.data_as(f32p)
    c = Counts()
    args = (gp, nx, ny, nz, sub, sign, x_origin, int(emit_last_plane), flags, vdiv, vmul, vadd)
    rc = lib.smb_oracle_mc(*args, None, None, ctypes.byref(c))
    if rc != 0:
        raise MemoryError("mc oracle allocation failed")
    verts = np.empty((c.nverts, 3), dtype=np.float32)
    faces = np.empty((c.ntris, 3), dtype=np.int64)
    rc = lib.smb_oracle_mc(
        *args, verts.ctypes.data_as(f32p), faces.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.byref(c)
    )
    if rc != 0:
        raise MemoryError("mc oracle allocation failed")
    return verts, faces, c


def cube_cases(grid: np.ndarray, sub: float = 0.0, sign: float = 1.0) -> np.ndarray:
    lib = _load()
    g = np.ascontiguousarray(grid, dtype=np.float32)
    nx, ny, nz = g.shape
    out = np.empty((nx - 1, ny - 1, nz - 1), dtype=np.uint8)
    lib.smb_oracle_mc_cases(
        g.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), nx, ny, nz, sub, sign,
        out.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
    )
    return out


def marching_cubes(volume: np.ndarray, level: float = 0.0):
    """Stand-in with skimage.measure.marching_cubes' call shape (isosurface.py:46-48):
    returns (verts, faces, normals, values); verts in index units / array-axis order;
    raises ValueError / RuntimeError like skimage does [memory, SURVEY 8b]."""
    v = np.ascontiguousarray(volume, dtype=np.float32)
    verts, faces, c = marching_cubes_slab(v, sub=float(level), sign=1.0)
    n = v.size
    if c.npos == n or c.nneg == n:
        raise ValueError("Surface level must be within volume data range.")
    if c.nverts == 0 or c.ntris == 0:
        raise RuntimeError("No surface found at the given iso value.")
    return verts, faces.astype(np.int32), None, None


def helper_forward(level_in: np.ndarray, resolution: int) -> Tuple[np.ndarray, np.ndarray]:
    """MarchingCubeHelper.forward (isosurface.py:41-54) on top of the oracle MC."""
    R = resolution
    g = np.ascontiguousarray(level_in, dtype=np.float32).reshape(R, R, R)
    verts, faces, c = marching_cubes_slab(g, sub=0.0, sign=-1.0, flags=FLIP | DIV, vdiv=float(R - 1.0))
    n = g.size
    if c.npos == n or c.nneg == n:
        raise ValueError("Surface level must be within volume data range.")
    if c.nverts == 0 or c.ntris == 0:
        raise RuntimeError("No surface found at the given iso value.")
    return verts, faces
