"""ctypes binding of ``include/sculptmate_b200.h`` (the C ABI of the CUDA library).

This is the stub a maintainer of the reference would add.
There is no CPU fallback: if the shared library is missing, ``load()`` raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_ubyte, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsculptmate_b200.so")

OK = 0
ERR_CUDA, ERR_BAD_ARG, ERR_WORKSPACE, ERR_ARCH, ERR_LEVEL_RANGE, ERR_NO_SURFACE, ERR_CAPACITY = -1, -2, -3, -4, -5, -6, -7
MC_FLIP, MC_DIV, MC_AFFINE, MC_FACES_I32, MC_COALESCE = 1, 2, 4, 8, 16


class DecoderLayout(ctypes.Structure):
    _fields_ = [
        ("total_bytes", c_uint32),
        ("n_hidden", c_uint32),
        ("off_tc_hidden", c_uint32),
        ("off_tc_final", c_uint32),
        ("off_tc_l0", c_uint32),
        ("off_bias_half", c_uint32),
        ("off_bias_final", c_uint32),
        ("off_w0_half", c_uint32),
        ("off_f32", c_uint32),
        ("off_tc_biasblk", c_uint32),
        ("reserved", c_uint32 * 6),
    ]


class QueryCfg(ctypes.Structure):
    _fields_ = [
        ("radius", c_float),
        ("density_bias", c_float),
        ("align_corners", c_int),
        ("Hp", c_int),
        ("Wp", c_int),
    ]


class MlpTcLayout(ctypes.Structure):
    _fields_ = [
        ("total_bytes", c_uint32),
        ("n_layers", c_uint32),
        ("kblocks", c_uint32 * 12),
        ("n_out", c_uint32 * 12),
        ("w_off", c_uint32 * 12),
        ("b_off", c_uint32 * 12),
    ]


class McCounts(ctypes.Structure):
    _fields_ = [
        ("nverts", c_int64),
        ("ntris", c_int64),
        ("nverts_numbered", c_int64),
        ("reserved", c_int64),
    ]


# every symbol include/sculptmate_b200.h declares: name -> (restype, argtypes)
_FLOATPP = POINTER(POINTER(c_float))
SIGNATURES = {
    "smb_status_string": (c_char_p, [c_int]),
    "smb_device_check": (c_int, []),
    "smb_version": (c_int, []),
    "smb_decoder_layout_for": (c_int, [c_int, POINTER(DecoderLayout)]),
    "smb_decoder_pack_host": (c_int, [_FLOATPP, _FLOATPP, c_int, POINTER(DecoderLayout), c_void_p]),
    "smb_scene_prepare": (c_int, [c_void_p, c_int, c_int, c_void_p, POINTER(DecoderLayout), c_void_p, c_void_p, c_void_p]),
    "smb_query_points_f32": (
        c_int,
        [c_void_p, c_void_p, POINTER(DecoderLayout), POINTER(QueryCfg), c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "smb_decoder_forward_f32": (c_int, [c_void_p, POINTER(DecoderLayout), c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "smb_lattice_axis_host": (c_int, [c_int, c_float, POINTER(c_float)]),
    "smb_query_lattice_tc": (
        c_int,
        [c_void_p, c_void_p, POINTER(DecoderLayout), POINTER(QueryCfg), c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    ),
    "smb_bake_workspace_bytes": (c_size_t, [c_int]),
    "smb_bake_rasterize": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "smb_bake_interpolate": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "smb_ray_sample_positions": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "smb_ray_composite": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "smb_mesh_loop_colors": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p]),
    "smb_mesh_faces_i32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "smb_mesh_faces_i64": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "smb_query_lattice_tc_signs": (
        c_int,
        [c_void_p, c_void_p, POINTER(DecoderLayout), POINTER(QueryCfg), c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_float,
         c_void_p, c_size_t, c_void_p],
    ),
    "smb_mc_count_presigned": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "smb_query_lattice_f32": (
        c_int,
        [c_void_p, c_void_p, POINTER(DecoderLayout), POINTER(QueryCfg), c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    ),
    "smb_query_tetgrid_tc": (
        c_int,
        [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "smb_mc_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "smb_mc_count": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "smb_mc_emit": (
        c_int,
        [c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_int, c_float, c_float, c_float, c_int64, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "smb_mc_emit_bounded": (
        c_int,
        [c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_int, c_float, c_float, c_float, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p],
    ),
    "smb_mc_emit_gather": (
        c_int,
        [c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p, c_int,
         c_void_p, c_int64, c_void_p, c_int64, c_void_p],
    ),
    "smb_peer_wait_release": (c_int, [c_void_p, c_int64, c_void_p]),
    "smb_peer_publish_counts": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p]),
    "smb_mc_emit_gather_flags": (
        c_int,
        [c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p, c_int64, c_int,
         c_void_p, c_int64, c_void_p, c_int64, c_void_p],
    ),
    "smb_peer_signal_done": (c_int, [c_void_p, c_int, c_int64, c_void_p]),
    "smb_peer_wait_all": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p]),
    "smb_dev_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "smb_dev_free": (c_int, [c_void_p]),
    "smb_ipc_export": (c_int, [c_void_p, c_void_p]),
    "smb_ipc_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "smb_ipc_close": (c_int, [c_void_p]),
    "smb_mc_cases": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p]),
    "smb_grid_minmax": (c_int, [c_void_p, c_int64, c_float, c_float, c_void_p, c_void_p]),
    "smb_extractor_create": (c_int, [_FLOATPP, _FLOATPP, c_int, c_float, c_float, c_int, c_int, POINTER(c_void_p)]),
    "smb_extractor_destroy": (None, [c_void_p]),
    "smb_extractor_set_axis": (c_int, [c_void_p, c_int, POINTER(c_float)]),
    "smb_extract_mesh_host": (
        c_int,
        [c_void_p, POINTER(c_float), c_int, c_float, POINTER(POINTER(c_float)), POINTER(POINTER(c_int64)), POINTER(c_int64), POINTER(c_int64)],
    ),
    "smb_extractor_pinned_input": (c_int, [c_void_p, POINTER(POINTER(c_float))]),
    "smb_extractor_set_faces_i32": (c_int, [c_void_p, c_int]),
    "smb_extract_mesh_device_to_host": (
        c_int,
        [c_void_p, c_void_p, c_int, c_float, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, POINTER(c_int64), POINTER(c_int64)],
    ),
    "smb_extractor_enable_timing": (c_int, [c_void_p, c_int]),
    "smb_extractor_last_timing": (c_int, [c_void_p, POINTER(c_float), POINTER(c_float), POINTER(c_float)]),
    "smb_extract_mesh_device": (
        c_int,
        [c_void_p, c_void_p, c_int, c_float, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_void_p, POINTER(c_int64), POINTER(c_int64)],
    ),
    "smb_extract_mesh_host_textured": (
        c_int,
        [c_void_p, POINTER(c_float), c_int, c_float, POINTER(POINTER(c_float)), POINTER(POINTER(c_int64)), POINTER(POINTER(c_float)),
         POINTER(POINTER(c_float)), POINTER(c_int64), POINTER(c_int64)],
    ),
    "smb_mlp_tc_layout_for": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(MlpTcLayout)]),
    "smb_mlp_tc_pack_host": (c_int, [_FLOATPP, _FLOATPP, POINTER(c_int), POINTER(c_int), POINTER(MlpTcLayout), c_void_p]),
    "smb_scene_prepare_half": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "smb_query_points_tc": (
        c_int,
        [c_void_p, c_int, c_int, c_int, c_int, c_void_p, POINTER(MlpTcLayout), c_float, c_float, c_int, c_void_p, c_int64, c_void_p, c_void_p,
         c_void_p, c_void_p, c_void_p],
    ),
    "smb_sf3d_heads_floats": (c_int, []),
    "smb_sf3d_query_f32": (
        c_int,
        [c_void_p, c_int, c_int, c_void_p, c_float, c_float, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "smb_mtet_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "smb_mtet_count": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_size_t, c_void_p, c_void_p]),
    "smb_mtet_emit": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "smb_mtet_emit_affine": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "smb_mtet_deform": (c_int, [c_void_p, c_void_p, c_float, c_int64, c_void_p, c_void_p]),
}

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the CUDA library.  Raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} is missing: build it with `python -m sculptmate_b200.build` "
                "(or __graft_entry__.build()). There is no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def status_string(rc: int) -> str:
    return load().smb_status_string(rc).decode()


class SmbError(RuntimeError):
    def __init__(self, rc: int, where: str):
        self.rc = rc
        super().__init__(f"{where}: {status_string(rc)} (status {rc})")


def check(rc: int, where: str) -> None:
    if rc != OK:
        raise SmbError(rc, where)
