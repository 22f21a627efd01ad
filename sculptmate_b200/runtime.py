"""Tensor-level host wrappers over the C ABI.

PyTorch is plumbing here: it owns device memory and streams; all arithmetic on
the hot path happens in the CUDA library.  Every function requires CUDA tensors
and raises otherwise -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import weakref
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _capi
from ._capi import DecoderLayout, McCounts, MlpTcLayout, QueryCfg, check

PLANE_CHANNELS = 40
HIDDEN = 64


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"sculptmate_b200: `{name}` must be a CUDA tensor (got {t.device}); the B200 path has no CPU fallback"
        )


def _stream_ptr(device: torch.device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else int(t.data_ptr())


# ------------------------------------------------------------------ decoder
def decoder_layout(n_hidden: int) -> DecoderLayout:
    lay = DecoderLayout()
    check(_capi.load().smb_decoder_layout_for(int(n_hidden), ctypes.byref(lay)), "smb_decoder_layout_for")
    return lay


def pack_decoder_host(weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, DecoderLayout]:
    """Pack NeRFMLP parameters (network_utils.py:48-79) into the kernel blob (host, uint8)."""
    n_hidden = len(weights) - 1
    if len(biases) != len(weights) or n_hidden < 2:
        raise ValueError("expected n_hidden+1 weight and bias tensors (n_hidden >= 2)")
    exp_shapes = [(HIDDEN, 3 * PLANE_CHANNELS)] + [(HIDDEN, HIDDEN)] * (n_hidden - 1) + [(4, HIDDEN)]
    ws, bs = [], []
    for l, (w, b) in enumerate(zip(weights, biases)):
        if tuple(w.shape) != exp_shapes[l] or tuple(b.shape) != (exp_shapes[l][0],):
            raise NotImplementedError(
                f"decoder layer {l} has shape {tuple(w.shape)}; the CUDA path is built for the TripoSR decoder "
                f"(in 120, width 64, out 4; config.yaml:25-30), expected {exp_shapes[l]}"
            )
        ws.append(w.detach().to(device="cpu", dtype=torch.float32).contiguous())
        bs.append(b.detach().to(device="cpu", dtype=torch.float32).contiguous())
    lay = decoder_layout(n_hidden)
    blob = torch.zeros(lay.total_bytes, dtype=torch.uint8)
    fpp = ctypes.POINTER(ctypes.c_float)
    W = (fpp * len(ws))(*[ctypes.cast(w.data_ptr(), fpp) for w in ws])
    B = (fpp * len(bs))(*[ctypes.cast(b.data_ptr(), fpp) for b in bs])
    check(
        _capi.load().smb_decoder_pack_host(W, B, n_hidden, ctypes.byref(lay), ctypes.c_void_p(blob.data_ptr())),
        "smb_decoder_pack_host",
    )
    return blob, lay


@dataclass
class DecoderPack:
    blob: torch.Tensor  # uint8 on the device
    layout: DecoderLayout
    key: Tuple

    @property
    def device(self) -> torch.device:
        return self.blob.device


def decoder_params(decoder: torch.nn.Module) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """Linear layers of a NeRFMLP in order (state-dict keys layers.{0,2,..}.{weight,bias})."""
    lin = [m for m in decoder.layers if isinstance(m, torch.nn.Linear)]
    if any(m.bias is None for m in lin):
        raise NotImplementedError("the CUDA path expects NeRFMLP with bias=True (config default)")
    acts = [m for m in decoder.layers if not isinstance(m, torch.nn.Linear)]
    if not all(isinstance(a, torch.nn.SiLU) for a in acts):
        raise NotImplementedError("the CUDA path implements activation='silu' (TripoSR config.yaml:29)")
    return [m.weight for m in lin], [m.bias for m in lin]


# keyed weakly by the decoder module: an entry dies with its decoder (no growth, no stale hit on a recycled id())
_pack_cache: "weakref.WeakKeyDictionary[torch.nn.Module, DecoderPack]" = weakref.WeakKeyDictionary()


def get_decoder_pack(decoder: torch.nn.Module, device: torch.device) -> DecoderPack:
    """Device blob for `decoder`, rebuilt when any parameter changes (ptr/version) or moves."""
    ws, bs = decoder_params(decoder)
    key = tuple((p.data_ptr(), p._version, str(p.device)) for p in (*ws, *bs)) + (str(device),)
    cached = _pack_cache.get(decoder)
    if cached is not None and cached.key == key:
        return cached
    blob, lay = pack_decoder_host(ws, bs)
    pack = DecoderPack(blob=blob.to(device), layout=lay, key=key)
    _pack_cache[decoder] = pack
    return pack


# ------------------------------------------------------------- scene planes
@dataclass
class ScenePlanes:
    planes_cl: Optional[torch.Tensor]  # (3,H,W,40) fp32
    planes_q: Optional[torch.Tensor]  # (3,H,W,64) fp32
    Hp: int
    Wp: int
    planes_h: Optional[torch.Tensor] = None  # (3,H,W,40) fp16, for the tensor-core points kernel


def prepare_scene(triplane: torch.Tensor, pack: DecoderPack, want_cl: bool = True, want_q: bool = True) -> ScenePlanes:
    _require_cuda(triplane, "triplane")
    if triplane.dim() != 4 or triplane.shape[0] != 3 or triplane.shape[1] != PLANE_CHANNELS:
        raise NotImplementedError(f"triplane must be (3,{PLANE_CHANNELS},Hp,Wp); got {tuple(triplane.shape)}")
    tp = triplane.detach().to(torch.float32).contiguous()
    _, _, Hp, Wp = tp.shape
    dev = tp.device
    cl = torch.empty((3, Hp, Wp, PLANE_CHANNELS), dtype=torch.float32, device=dev) if want_cl else None
    q = torch.empty((3, Hp, Wp, HIDDEN), dtype=torch.float32, device=dev) if want_q else None
    with torch.cuda.device(dev):
        check(
            _capi.load().smb_scene_prepare(
                tp.data_ptr(), Hp, Wp, pack.blob.data_ptr(), ctypes.byref(pack.layout), _ptr(cl), _ptr(q), _stream_ptr(dev)
            ),
            "smb_scene_prepare",
        )
    return ScenePlanes(cl, q, Hp, Wp)


def _cfg(radius: float, density_bias: float, Hp: int, Wp: int, align_corners: bool = False) -> QueryCfg:
    return QueryCfg(float(radius), float(density_bias), int(bool(align_corners)), int(Hp), int(Wp))


# -------------------------------------------------------------- field query
def query_points(
    planes: ScenePlanes,
    pack: DecoderPack,
    positions: torch.Tensor,
    radius: float,
    density_bias: float,
    want: Sequence[str] = ("density", "features", "density_act", "color"),
    align_corners: bool = False,
) -> Dict[str, torch.Tensor]:
    """query_triplane for arbitrary positions (n,3) in (-radius, radius); fp32 CUDA-core kernel."""
    _require_cuda(positions, "positions")
    pos = positions.detach().to(torch.float32).contiguous().view(-1, 3)
    n = pos.shape[0]
    dev = pos.device
    widths = {"density": 1, "features": 3, "density_act": 1, "color": 3}
    outs = {k: torch.empty((n, widths[k]), dtype=torch.float32, device=dev) for k in want}
    cfg = _cfg(radius, density_bias, planes.Hp, planes.Wp, align_corners)
    with torch.cuda.device(dev):
        check(
            _capi.load().smb_query_points_f32(
                planes.planes_cl.data_ptr(), pack.blob.data_ptr(), ctypes.byref(pack.layout), ctypes.byref(cfg),
                pos.data_ptr(), n, _ptr(outs.get("density")), _ptr(outs.get("features")),
                _ptr(outs.get("density_act")), _ptr(outs.get("color")), _stream_ptr(dev),
            ),
            "smb_query_points_f32",
        )
    return outs


def decoder_forward(pack: DecoderPack, features: torch.Tensor) -> Dict[str, torch.Tensor]:
    """NeRFMLP.forward on (n,120) features -> {"density": (n,1), "features": (n,3)} (fp32 CUDA kernel)."""
    _require_cuda(features, "x")
    x = features.detach().to(torch.float32).contiguous().view(-1, 3 * PLANE_CHANNELS)
    n, dev = x.shape[0], x.device
    d = torch.empty((n, 1), dtype=torch.float32, device=dev)
    f = torch.empty((n, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_capi.load().smb_decoder_forward_f32(pack.blob.data_ptr(), ctypes.byref(pack.layout), x.data_ptr(), n, d.data_ptr(), f.data_ptr(), _stream_ptr(dev)), "smb_decoder_forward_f32")
    return {"density": d, "features": f}


def lattice_axis(resolution: int, radius: float, points_range=(0, 1), device=None) -> torch.Tensor:
    """Per-axis lattice coordinate mapped to (-1,1) with the reference's own torch ops:
    linspace (isosurface.py:30-32) -> scale_tensor (system.py:177-181) -> scale_tensor
    (nerf_renderer.py:52-54).  The lattice is separable, so R values describe all R^3 rows."""
    from .tsr.utils import scale_tensor

    a = torch.linspace(*points_range, resolution)
    a = scale_tensor(a, points_range, (-radius, radius))
    a = scale_tensor(a, (-radius, radius), (-1, 1))
    return a.to(device) if device is not None else a


def query_lattice(
    planes: ScenePlanes,
    pack: DecoderPack,
    axis_u: torch.Tensor,
    resolution: int,
    radius: float,
    density_bias: float,
    x_begin: int = 0,
    nx: Optional[int] = None,
    precision: str = "tc",
    want_raw: bool = False,
    out: Optional[torch.Tensor] = None,
    mc_signs: Optional[Tuple[float, float]] = None,
):
    """density_act on x-planes [x_begin, x_begin+nx) of the R^3 lattice -> (nx,R,R) fp32.

    ``mc_signs=(sub, sign)`` (tensor-core path only): the kernel also writes the marching-cubes sign
    masks of ``(density_act - sub) * sign > 0`` into the cached MC workspace of this slab shape, so that
    ``mc_extract(..., presigned=True)`` on the returned grid skips the classification pass."""
    R = int(resolution)
    nx = R - x_begin if nx is None else int(nx)
    _require_cuda(axis_u, "axis_u")
    dev = axis_u.device
    if out is None:
        out = torch.empty((nx, R, R), dtype=torch.float32, device=dev)
    raw = torch.empty((nx, R, R), dtype=torch.float32, device=dev) if want_raw else None
    cfg = _cfg(radius, density_bias, planes.Hp, planes.Wp, False)
    lib = _capi.load()
    with torch.cuda.device(dev):
        if precision == "tc" and mc_signs is not None:
            w = _mc_cache.get(dev, (nx, R, R))
            w.generation += 1  # records of an earlier count pass on this workspace no longer match its sign masks
            rc = lib.smb_query_lattice_tc_signs(
                planes.planes_q.data_ptr(), pack.blob.data_ptr(), ctypes.byref(pack.layout), ctypes.byref(cfg),
                axis_u.data_ptr(), R, int(x_begin), nx, out.data_ptr(), _ptr(raw), float(mc_signs[0]), float(mc_signs[1]),
                w.ws.data_ptr(), w.ws.numel(), _stream_ptr(dev),
            )
            check(rc, "smb_query_lattice_tc_signs")
            w.mark_signed(out, mc_signs[0], mc_signs[1])
        elif precision == "tc":
            rc = lib.smb_query_lattice_tc(
                planes.planes_q.data_ptr(), pack.blob.data_ptr(), ctypes.byref(pack.layout), ctypes.byref(cfg),
                axis_u.data_ptr(), R, int(x_begin), nx, out.data_ptr(), _ptr(raw), _stream_ptr(dev),
            )
            check(rc, "smb_query_lattice_tc")
        elif precision == "fp32":
            if mc_signs is not None:
                raise ValueError("mc_signs needs precision='tc'")
            rc = lib.smb_query_lattice_f32(
                planes.planes_cl.data_ptr(), pack.blob.data_ptr(), ctypes.byref(pack.layout), ctypes.byref(cfg),
                axis_u.data_ptr(), R, int(x_begin), nx, out.data_ptr(), _ptr(raw), _stream_ptr(dev),
            )
            check(rc, "smb_query_lattice_f32")
        else:
            raise ValueError(f"precision must be 'tc' or 'fp32', got {precision!r}")
    return (out, raw) if want_raw else out


# ---------------------------------------------------------- marching cubes
class McWorkspace:
    """Scratch of one (device, slab shape): the word records / sign masks, the device and pinned count records.

    ``generation`` counts the count passes run on it and ``signed_for`` remembers which density tensor the lattice
    kernel balloted the sign masks for: a pending emit (or a ``presigned`` count) is only valid against the very
    pass / grid it belongs to, otherwise it would read another grid's records (same shape, same cache entry)."""

    def __init__(self, device: torch.device, shape: Tuple[int, int, int]) -> None:
        nbytes = _capi.load().smb_mc_workspace_bytes(*shape)
        self.shape = tuple(shape)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.counts_dev = torch.zeros(4, dtype=torch.int64, device=device)
        self.counts_pin = torch.zeros(4, dtype=torch.int64).pin_memory()
        self.generation = 0
        self.signed_for = None  # (weakref to the grid tensor, grid._version, sub, sign)

    def mark_signed(self, grid: torch.Tensor, sub: float, sign: float) -> None:
        self.signed_for = (weakref.ref(grid), grid._version, float(sub), float(sign))

    def signed_matches(self, grid: torch.Tensor, sub: float, sign: float) -> bool:
        sf = self.signed_for
        return sf is not None and sf[0]() is grid and sf[1] == grid._version and sf[2] == float(sub) and sf[3] == float(sign)


class McWorkspaceCache:
    """Per-(device, shape) workspaces, cached like the reference caches its helper per resolution
    (system.py:118-124).  Least-recently-used eviction; a workspace that a live ``McPending`` still references
    stays alive through that reference."""

    MAX_ENTRIES = 6

    def __init__(self) -> None:
        self._ws: "Dict[Tuple, McWorkspace]" = {}

    def get(self, device: torch.device, shape: Tuple[int, int, int]) -> McWorkspace:
        key = (str(device), tuple(shape))
        w = self._ws.pop(key, None)
        if w is None:
            w = McWorkspace(device, shape)
            while len(self._ws) >= self.MAX_ENTRIES:
                self._ws.pop(next(iter(self._ws)))
        self._ws[key] = w  # most recently used last
        return w


_mc_cache = McWorkspaceCache()


@dataclass
class McPending:
    grid: torch.Tensor
    sub: float
    sign: float
    emit_last_plane: bool
    nverts: int
    ntris: int
    nverts_numbered: int
    wsp: Optional[McWorkspace] = None  # the workspace that holds this pass' records
    generation: int = -1

    def stale(self) -> bool:
        return self.wsp is None or self.wsp.generation != self.generation


def _launch_count(lib, grid, nx, ny, nz, sub, sign, emit_last_plane, w: McWorkspace, st, presigned: bool) -> None:
    ws, counts_dev = w.ws, w.counts_dev
    w.generation += 1
    if presigned and w.signed_matches(grid, sub, sign):
        # the sign masks in ws were written by query_lattice(mc_signs=(sub, sign)) for this very grid
        check(lib.smb_mc_count_presigned(nx, ny, nz, int(emit_last_plane), ws.data_ptr(), ws.numel(), counts_dev.data_ptr(), st), "smb_mc_count_presigned")
    else:
        w.signed_for = None  # the stand-alone sign pass overwrites the masks
        check(lib.smb_mc_count(grid.data_ptr(), nx, ny, nz, float(sub), float(sign), int(emit_last_plane), ws.data_ptr(), ws.numel(),
                               counts_dev.data_ptr(), st), "smb_mc_count")


def _check_grid(grid: torch.Tensor) -> None:
    _require_cuda(grid, "grid")
    if grid.dim() != 3 or grid.dtype != torch.float32 or not grid.is_contiguous():
        raise ValueError("grid must be a contiguous (nx,ny,nz) float32 tensor")


def mc_count(grid: torch.Tensor, sub: float = 0.0, sign: float = 1.0, emit_last_plane: bool = True, presigned: bool = False) -> McPending:
    """Classify + scan; returns the counts (one host sync to read them)."""
    _check_grid(grid)
    nx, ny, nz = grid.shape
    dev = grid.device
    w = _mc_cache.get(dev, (nx, ny, nz))
    with torch.cuda.device(dev):
        _launch_count(_capi.load(), grid, nx, ny, nz, sub, sign, emit_last_plane, w, _stream_ptr(dev), presigned)
        w.counts_pin.copy_(w.counts_dev, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    c = w.counts_pin
    return McPending(grid, float(sub), float(sign), bool(emit_last_plane), int(c[0]), int(c[1]), int(c[2]), w, w.generation)


def _faces_dtype(flags: int) -> torch.dtype:
    return torch.int32 if flags & _capi.MC_FACES_I32 else torch.int64


def mc_emit(
    pending: McPending,
    x_origin: int = 0,
    flags: int = 0,
    vdiv: float = 1.0,
    vmul: float = 1.0,
    vadd: float = 0.0,
    vertex_id_offset: int = 0,
    verts_out: Optional[torch.Tensor] = None,
    faces_out: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    grid = pending.grid
    nx, ny, nz = grid.shape
    dev = grid.device
    if pending.stale():
        # another count pass has reused the workspace since (same shape): redo this grid's pass, do not read foreign records
        redo = mc_count(grid, pending.sub, pending.sign, pending.emit_last_plane)
        if (redo.nverts, redo.ntris) != (pending.nverts, pending.ntris):
            raise RuntimeError("the density grid changed between mc_count and mc_emit")
        pending.wsp, pending.generation = redo.wsp, redo.generation
    fdt = _faces_dtype(flags)
    verts = verts_out if verts_out is not None else torch.empty((pending.nverts, 3), dtype=torch.float32, device=dev)
    faces = faces_out if faces_out is not None else torch.empty((pending.ntris, 3), dtype=fdt, device=dev)
    if faces.dtype != fdt:
        raise ValueError(f"faces_out must be {fdt} for these flags")
    if fdt == torch.int32 and vertex_id_offset + pending.nverts_numbered >= 2**31:
        raise ValueError("vertex ids do not fit int32; drop MC_FACES_I32")
    if pending.nverts == 0 and pending.ntris == 0:
        return verts, faces
    with torch.cuda.device(dev):
        check(
            _capi.load().smb_mc_emit_bounded(
                grid.data_ptr(), nx, ny, nz, pending.sub, pending.sign, int(x_origin), int(pending.emit_last_plane),
                int(flags), float(vdiv), float(vmul), float(vadd), int(vertex_id_offset), pending.wsp.ws.data_ptr(),
                verts.data_ptr(), int(verts.shape[0]), faces.data_ptr(), int(faces.shape[0]), _stream_ptr(dev),
            ),
            "smb_mc_emit_bounded",
        )
    return verts, faces


_mc_caps: Dict[Tuple, Tuple[int, int]] = {}


def mc_extract(
    grid: torch.Tensor, sub: float = 0.0, sign: float = 1.0, flags: int = 0, vdiv: float = 1.0, vmul: float = 1.0, vadd: float = 0.0,
    presigned: bool = False, on_launched=None,
) -> Tuple[torch.Tensor, torch.Tensor, McPending]:
    """count + emit for a whole grid with no host round trip between them: emit is launched right
    behind count into buffers sized from the previous mesh of this shape (+25 %), the counts are
    read afterwards and the outputs are views of exactly (V,3) / (F,3).  Falls back to the
    two-phase path on the first call for a shape and on overflow.  ``presigned`` is honoured only when the
    workspace's sign masks were balloted for this very grid (``query_lattice(mc_signs=...)``)."""
    _check_grid(grid)
    nx, ny, nz = grid.shape
    dev = grid.device
    key = (str(dev), (nx, ny, nz))
    cap = _mc_caps.get(key)
    fdt = _faces_dtype(flags)
    if cap is None:
        pend = mc_count(grid, sub=sub, sign=sign, emit_last_plane=True, presigned=presigned)
        verts, faces = mc_emit(pend, flags=flags, vdiv=vdiv, vmul=vmul, vadd=vadd)
        if on_launched is not None:
            on_launched()
    else:
        w = _mc_cache.get(dev, (nx, ny, nz))
        lib = _capi.load()
        verts = torch.empty((cap[0], 3), dtype=torch.float32, device=dev)
        faces = torch.empty((cap[1], 3), dtype=fdt, device=dev)
        with torch.cuda.device(dev):
            st = _stream_ptr(dev)
            _launch_count(lib, grid, nx, ny, nz, sub, sign, True, w, st, presigned)
            check(
                lib.smb_mc_emit_bounded(grid.data_ptr(), nx, ny, nz, float(sub), float(sign), 0, 1, int(flags), float(vdiv), float(vmul),
                                        float(vadd), 0, w.ws.data_ptr(), verts.data_ptr(), cap[0], faces.data_ptr(), cap[1], st),
                "smb_mc_emit_bounded",
            )
            if on_launched is not None:  # e.g. a CUDA event: the kernels are queued, the host has not synchronised yet
                on_launched()
            w.counts_pin.copy_(w.counts_dev, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        c = w.counts_pin
        pend = McPending(grid, float(sub), float(sign), True, int(c[0]), int(c[1]), int(c[2]), w, w.generation)
        if pend.nverts <= cap[0] and pend.ntris <= cap[1]:
            verts, faces = verts[: pend.nverts], faces[: pend.ntris]
        else:  # the surface grew past the remembered capacity: emit again at the exact size
            verts, faces = mc_emit(pend, flags=flags, vdiv=vdiv, vmul=vmul, vadd=vadd)
    old = cap or (0, 0)
    _mc_caps[key] = (max(old[0], pend.nverts * 5 // 4 + 1024), max(old[1], pend.ntris * 5 // 4 + 1024))
    if len(_mc_caps) > 8:
        for k in list(_mc_caps)[:-8]:
            del _mc_caps[k]
    return verts, faces, pend


# ------------------------------------------------------ whole path, one C call
class MeshExtractor:
    """``smb_extractor`` handle for one decoder on one device: the device-resident whole path of
    TSR.extract_mesh (tsr/system.py:173-189) as ONE C call per scene code (``smb_extract_mesh_device``) --
    prepare, lattice query with the sign ballot, count, totals and emit are queued back to back by the library,
    so no Python runs between the launches.  Output tensors are allocated here per call (the caller owns them);
    their capacity is remembered per resolution from the previous mesh (+25 %), and an overflow repeats only
    the emit."""

    def __init__(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], radius: float, density_bias: float,
                 Hp: int, Wp: int, device: torch.device) -> None:
        lib = _capi.load()
        ws = [w.detach().to(device="cpu", dtype=torch.float32).contiguous() for w in weights]
        bs = [b.detach().to(device="cpu", dtype=torch.float32).contiguous() for b in biases]
        fpp = ctypes.POINTER(ctypes.c_float)
        W = (fpp * len(ws))(*[ctypes.cast(w.data_ptr(), fpp) for w in ws])
        B = (fpp * len(bs))(*[ctypes.cast(b.data_ptr(), fpp) for b in bs])
        self.device = device
        self.radius = float(radius)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(device):
            check(lib.smb_extractor_create(W, B, len(ws) - 1, float(radius), float(density_bias), int(Hp), int(Wp), ctypes.byref(self._h)), "smb_extractor_create")
        self._axis_set = set()
        self._caps: Dict[int, Tuple[int, int]] = {}
        self._host_caps: Dict[int, Tuple[int, int]] = {}

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            try:
                _capi.load().smb_extractor_destroy(self._h)
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass
            self._h = None

    __del__ = close

    def enable_timing(self, on: bool = True) -> None:
        """CUDA events around prepare / lattice kernel / marching cubes inside ``extract`` (see ``last_timing``)."""
        check(_capi.load().smb_extractor_enable_timing(self._h, int(on)), "smb_extractor_enable_timing")

    def last_timing(self) -> Tuple[float, float, float]:
        """(prepare_ms, lattice_ms, mc_ms) of the last ``extract`` call (mc = count + totals + emit on the device)."""
        a, b, c = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
        check(_capi.load().smb_extractor_last_timing(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "smb_extractor_last_timing")
        return a.value, b.value, c.value

    def extract(self, triplane: torch.Tensor, resolution: int, threshold: float, faces_dtype: torch.dtype = torch.int64,
                axis_u: Optional[torch.Tensor] = None, want_density: bool = False):
        """-> (v_pos (V,3) fp32 in (-radius, radius), t_pos_idx (F,3) ``faces_dtype``[, density_act (R,R,R)]) on the device.
        Raises ValueError / RuntimeError like skimage for an iso level outside the data range / an empty surface."""
        _require_cuda(triplane, "triplane")
        if faces_dtype not in (torch.int64, torch.int32):
            raise ValueError("faces_dtype must be torch.int64 (the reference's LongTensor) or torch.int32")
        lib = _capi.load()
        dev = triplane.device
        R = int(resolution)
        tp = triplane.detach().to(torch.float32).contiguous()
        fpp = ctypes.POINTER(ctypes.c_float)
        with torch.cuda.device(dev):
            if axis_u is not None and R not in self._axis_set:
                # coordinates built with the reference's own torch ops (smb_lattice_axis_host restates them bit for bit;
                # passing them keeps that true on a host whose aten rounds differently)
                a = axis_u.detach().to("cpu", torch.float32).contiguous()
                check(lib.smb_extractor_set_axis(self._h, R, ctypes.cast(a.data_ptr(), fpp)), "smb_extractor_set_axis")
                self._axis_set = {R}
            vcap, fcap = self._caps.get(R, (0, 0))
            flags = _capi.MC_FACES_I32 if faces_dtype == torch.int32 else 0
            dens = torch.empty((R, R, R), dtype=torch.float32, device=dev) if want_density else None
            nv, nt = ctypes.c_int64(), ctypes.c_int64()
            emit_only = 0
            while True:
                verts = torch.empty((vcap, 3), dtype=torch.float32, device=dev)
                faces = torch.empty((fcap, 3), dtype=faces_dtype, device=dev)
                rc = lib.smb_extract_mesh_device(self._h, tp.data_ptr(), R, float(threshold), flags, _ptr(verts) if vcap else None, vcap,
                                                 _ptr(faces) if fcap else None, fcap, _ptr(dens), emit_only, _stream_ptr(dev),
                                                 ctypes.byref(nv), ctypes.byref(nt))
                if rc != _capi.ERR_CAPACITY:
                    break
                vcap, fcap, emit_only = max(vcap, nv.value * 5 // 4 + 1024), max(fcap, nt.value * 5 // 4 + 1024), 1
        if rc == _capi.ERR_LEVEL_RANGE:
            raise ValueError("Surface level must be within volume data range.")
        if rc == _capi.ERR_NO_SURFACE:
            raise RuntimeError("No surface found at the given iso value.")
        check(rc, "smb_extract_mesh_device")
        self._caps[R] = (max(vcap, nv.value * 5 // 4 + 1024), max(fcap, nt.value * 5 // 4 + 1024))
        if len(self._caps) > 4:
            self._caps.pop(next(iter(self._caps)))
        out = (verts[: nv.value], faces[: nt.value])
        return out + (dens,) if want_density else out


    def extract_to_host(self, triplane: torch.Tensor, resolution: int, threshold: float, faces_dtype: torch.dtype = torch.int64,
                        axis_u: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (v_pos, t_pos_idx) as PINNED HOST tensors the caller owns (fresh per call, from torch's caching pinned
        allocator): the slab pipeline of the library moves each slab's part of the mesh across PCIe while the next slab is
        computed, so only the tail of the device->host copy is exposed.  Same errors as ``extract``."""
        _require_cuda(triplane, "triplane")
        lib = _capi.load()
        dev = triplane.device
        R = int(resolution)
        tp = triplane.detach().to(torch.float32).contiguous()
        fpp = ctypes.POINTER(ctypes.c_float)
        with torch.cuda.device(dev):
            if axis_u is not None and R not in self._axis_set:
                a = axis_u.detach().to("cpu", torch.float32).contiguous()
                check(lib.smb_extractor_set_axis(self._h, R, ctypes.cast(a.data_ptr(), fpp)), "smb_extractor_set_axis")
                self._axis_set = {R}
            vcap, fcap = self._host_caps.get(R, (0, 0))
            flags = _capi.MC_FACES_I32 if faces_dtype == torch.int32 else 0
            nv, nt = ctypes.c_int64(), ctypes.c_int64()
            while True:
                hv = torch.empty((vcap, 3), dtype=torch.float32, pin_memory=True) if vcap else None
                hf = torch.empty((fcap, 3), dtype=faces_dtype, pin_memory=True) if fcap else None
                rc = lib.smb_extract_mesh_device_to_host(self._h, tp.data_ptr(), R, float(threshold), flags, _ptr(hv), vcap, _ptr(hf), fcap,
                                                         _stream_ptr(dev), ctypes.byref(nv), ctypes.byref(nt))
                if rc != _capi.ERR_CAPACITY:
                    break
                vcap, fcap = max(vcap, nv.value * 5 // 4 + 1024), max(fcap, nt.value * 5 // 4 + 1024)
        if rc == _capi.ERR_LEVEL_RANGE:
            raise ValueError("Surface level must be within volume data range.")
        if rc == _capi.ERR_NO_SURFACE:
            raise RuntimeError("No surface found at the given iso value.")
        check(rc, "smb_extract_mesh_device_to_host")
        self._host_caps[R] = (max(vcap, nv.value * 5 // 4 + 1024), max(fcap, nt.value * 5 // 4 + 1024))
        if len(self._host_caps) > 4:
            self._host_caps.pop(next(iter(self._host_caps)))
        return hv[: nv.value], hf[: nt.value]


_extractor_cache: "weakref.WeakKeyDictionary[torch.nn.Module, Dict[Tuple, MeshExtractor]]" = weakref.WeakKeyDictionary()


def get_mesh_extractor(decoder: torch.nn.Module, radius: float, density_bias: float, Hp: int, Wp: int, device: torch.device) -> MeshExtractor:
    """One MeshExtractor per (decoder parameters, plane size, device), rebuilt when a parameter changes."""
    ws, bs = decoder_params(decoder)
    key = _param_key((*ws, *bs), device) + (float(radius), float(density_bias), int(Hp), int(Wp))
    slot = _extractor_cache.setdefault(decoder, {})
    ex = slot.get("ex")
    if ex is None or slot.get("key") != key:
        if ex is not None:
            ex.close()
        ex = MeshExtractor(ws, bs, radius, density_bias, Hp, Wp, device)
        slot["ex"], slot["key"] = ex, key
    return ex


def mc_cases(grid: torch.Tensor, sub: float = 0.0, sign: float = 1.0) -> torch.Tensor:
    _require_cuda(grid, "grid")
    nx, ny, nz = grid.shape
    out = torch.empty((nx - 1, ny - 1, nz - 1), dtype=torch.uint8, device=grid.device)
    with torch.cuda.device(grid.device):
        check(
            _capi.load().smb_mc_cases(grid.data_ptr(), nx, ny, nz, float(sub), float(sign), out.data_ptr(), _stream_ptr(grid.device)),
            "smb_mc_cases",
        )
    return out


def grid_minmax(grid: torch.Tensor, sub: float = 0.0, sign: float = 1.0) -> Tuple[float, float]:
    _require_cuda(grid, "grid")
    out = torch.empty(2, dtype=torch.float32, device=grid.device)
    with torch.cuda.device(grid.device):
        check(
            _capi.load().smb_grid_minmax(grid.data_ptr(), grid.numel(), float(sub), float(sign), out.data_ptr(), _stream_ptr(grid.device)),
            "smb_grid_minmax",
        )
    lo, hi = out.cpu().tolist()
    return lo, hi


def raise_for_empty_surface(grid: torch.Tensor, sub: float, sign: float) -> None:
    """Same exception types skimage raises at isosurface.py:46-48."""
    lo, hi = grid_minmax(grid, sub, sign)
    if lo > 0.0 or hi < 0.0:
        raise ValueError("Surface level must be within volume data range.")
    raise RuntimeError("No surface found at the given iso value.")


# ------------------------------------------------------------------- SF3D
def prepare_planes_half(triplane: torch.Tensor) -> ScenePlanes:
    """Channels-last fp16 copy of a (3,40,Hp,Wp) triplane for the tensor-core points kernel."""
    _require_cuda(triplane, "triplane")
    if triplane.dim() != 4 or triplane.shape[0] != 3 or triplane.shape[1] != PLANE_CHANNELS:
        raise NotImplementedError(f"triplane must be (3,{PLANE_CHANNELS},Hp,Wp); got {tuple(triplane.shape)}")
    tp = triplane.detach().to(torch.float32).contiguous()
    _, _, Hp, Wp = tp.shape
    ph = torch.empty((3, Hp, Wp, PLANE_CHANNELS), dtype=torch.float16, device=tp.device)
    with torch.cuda.device(tp.device):
        check(_capi.load().smb_scene_prepare_half(tp.data_ptr(), Hp, Wp, ph.data_ptr(), _stream_ptr(tp.device)), "smb_scene_prepare_half")
    return ScenePlanes(None, None, Hp, Wp, planes_h=ph)


def prepare_planes_cl(triplane: torch.Tensor) -> ScenePlanes:
    """Channels-last copy of a (3,40,Hp,Wp) triplane (no decoder needed)."""
    _require_cuda(triplane, "triplane")
    if triplane.dim() != 4 or triplane.shape[0] != 3 or triplane.shape[1] != PLANE_CHANNELS:
        raise NotImplementedError(f"triplane must be (3,{PLANE_CHANNELS},Hp,Wp); got {tuple(triplane.shape)}")
    tp = triplane.detach().to(torch.float32).contiguous()
    _, _, Hp, Wp = tp.shape
    cl = torch.empty((3, Hp, Wp, PLANE_CHANNELS), dtype=torch.float32, device=tp.device)
    with torch.cuda.device(tp.device):
        check(_capi.load().smb_scene_prepare(tp.data_ptr(), Hp, Wp, None, None, cl.data_ptr(), None, _stream_ptr(tp.device)), "smb_scene_prepare")
    return ScenePlanes(cl, None, Hp, Wp)


# weak in the decoder module: an entry dies with its decoder (no growth, no stale hit on a recycled id())
_sf3d_heads_cache: "weakref.WeakKeyDictionary[torch.nn.Module, Tuple[Tuple, torch.Tensor]]" = weakref.WeakKeyDictionary()


def get_sf3d_heads(decoder: torch.nn.Module, device: torch.device) -> torch.Tensor:
    """MaterialMLP heads ``density`` and ``vertex_offset`` (network.py:158-178) packed into the
    fp32 device blob of include/sculptmate_b200.h; rebuilt when a parameter changes."""
    params = []
    for name in ("density", "vertex_offset"):
        seq = decoder.heads[name]
        lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
        if len(lin) != 3:
            raise NotImplementedError(f"head {name!r}: the CUDA path expects 2 hidden layers (config.yaml:50-65)")
        for m in lin:
            params += [m.weight, m.bias]
    key = tuple((p.data_ptr(), p._version, str(p.device)) for p in params) + (str(device),)
    hit = _sf3d_heads_cache.get(decoder)
    if hit is not None and hit[0] == key:
        return hit[1]
    total = _capi.load().smb_sf3d_heads_floats()
    blob = torch.zeros(total, dtype=torch.float32)
    off = 0
    for h in range(2):
        start = off
        for p_ in params[6 * h : 6 * h + 6]:
            flat = p_.detach().to(device="cpu", dtype=torch.float32).reshape(-1)
            blob[off : off + flat.numel()] = flat
            off += flat.numel()
        off = start + (off - start + 3) // 4 * 4  # each head is padded to a multiple of 4 floats
    assert off == total, (off, total)
    blob = blob.to(device)
    _sf3d_heads_cache[decoder] = (key, blob)
    return blob


def sf3d_query(
    planes: Optional[ScenePlanes],
    heads_blob: Optional[torch.Tensor],
    density_out_bias: float,
    radius: float,
    positions: Optional[torch.Tensor] = None,
    features: Optional[torch.Tensor] = None,
    want: Sequence[str] = ("density_act", "vertex_offset"),
) -> Dict[str, torch.Tensor]:
    """SF3D.query_triplane (+ the two MaterialMLP heads) for (n,3) positions, or the heads alone
    on (n,120) features.  ``want`` from {features, density_raw, density_act, vertex_offset}."""
    src = positions if positions is not None else features
    _require_cuda(src, "positions" if positions is not None else "features")
    src = src.detach().to(torch.float32).contiguous()
    n = src.shape[0]
    dev = src.device
    widths = {"features": 3 * PLANE_CHANNELS, "density_raw": 1, "density_act": 1, "vertex_offset": 3}
    outs = {k: torch.empty((n, widths[k]), dtype=torch.float32, device=dev) for k in want}
    with torch.cuda.device(dev):
        check(
            _capi.load().smb_sf3d_query_f32(
                _ptr(planes.planes_cl) if planes is not None else None, planes.Hp if planes is not None else 0,
                planes.Wp if planes is not None else 0, _ptr(heads_blob), float(radius), float(density_out_bias),
                src.data_ptr() if positions is not None else None, src.data_ptr() if positions is None else None, n,
                _ptr(outs.get("features")), _ptr(outs.get("density_raw")), _ptr(outs.get("density_act")),
                _ptr(outs.get("vertex_offset")), _stream_ptr(dev),
            ),
            "smb_sf3d_query_f32",
        )
    return outs


def mtet_deform(base: torch.Tensor, deform: torch.Tensor, scale: float) -> torch.Tensor:
    _require_cuda(deform, "deformation")
    base = base.to(deform.device, torch.float32).contiguous()
    deform = deform.contiguous()
    out = torch.empty_like(base)
    with torch.cuda.device(base.device):
        check(_capi.load().smb_mtet_deform(base.data_ptr(), deform.data_ptr(), float(scale), base.shape[0], out.data_ptr(), _stream_ptr(base.device)), "smb_mtet_deform")
    return out


_mtet_ws: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}


def marching_tets(
    positions: torch.Tensor, sdf: torch.Tensor, edges: torch.Tensor, tets: torch.Tensor, tet_edges: torch.Tensor,
    affine: Optional[Sequence[float]] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """MarchingTetrahedraHelper._forward on static int32 topology -> (verts (V,3) f32, faces (F,3) i64).
    ``affine`` (8 floats: in_lo, in_hi - in_lo, (out_hi - out_lo)[3], out_lo[3]) applies the caller's ``scale_tensor`` to the
    vertices inside the emit kernel (same fp32 operations, one pass instead of four elementwise kernels)."""
    _require_cuda(sdf, "level")
    dev = sdf.device
    ne, nt = int(edges.shape[0]), int(tets.shape[0])
    lib = _capi.load()
    key = (str(dev), ne, nt)
    if key not in _mtet_ws:
        if len(_mtet_ws) > 2:
            _mtet_ws.clear()
        _mtet_ws[key] = (
            torch.empty(lib.smb_mtet_workspace_bytes(ne, nt), dtype=torch.uint8, device=dev),
            torch.zeros(4, dtype=torch.int64, device=dev), torch.zeros(4, dtype=torch.int64).pin_memory(),
        )
    ws, counts_dev, counts_pin = _mtet_ws[key]
    positions = positions.contiguous()
    with torch.cuda.device(dev):
        check(lib.smb_mtet_count(sdf.data_ptr(), edges.data_ptr(), ne, tets.data_ptr(), nt, ws.data_ptr(), ws.numel(), counts_dev.data_ptr(), _stream_ptr(dev)), "smb_mtet_count")
        counts_pin.copy_(counts_dev, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        V, F = int(counts_pin[0]), int(counts_pin[1])
        verts = torch.empty((V, 3), dtype=torch.float32, device=dev)
        faces = torch.empty((F, 3), dtype=torch.int64, device=dev)
        if (V or F) and affine is None:
            check(
                lib.smb_mtet_emit(positions.data_ptr(), sdf.data_ptr(), edges.data_ptr(), ne, tet_edges.data_ptr(), nt, ws.data_ptr(),
                                  verts.data_ptr(), faces.data_ptr(), _stream_ptr(dev)),
                "smb_mtet_emit",
            )
        elif V or F:
            af = (ctypes.c_float * 8)(*[float(x) for x in affine])
            check(
                lib.smb_mtet_emit_affine(positions.data_ptr(), sdf.data_ptr(), edges.data_ptr(), ne, tet_edges.data_ptr(), nt, ws.data_ptr(),
                                         verts.data_ptr(), faces.data_ptr(), af, _stream_ptr(dev)),
                "smb_mtet_emit_affine",
            )
    return verts, faces


# ----------------------------------------------- tensor-core MLP on arbitrary positions
@dataclass
class MlpTcPack:
    blob: torch.Tensor  # uint8 on the device
    layout: MlpTcLayout
    key: Tuple


def _pack_mlp_tc(weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], device: torch.device, key: Tuple) -> MlpTcPack:
    ws = [w.detach().to(device="cpu", dtype=torch.float32).contiguous() for w in weights]
    bs = [b.detach().to(device="cpu", dtype=torch.float32).contiguous() for b in biases]
    n = len(ws)
    k_in = (ctypes.c_int * n)(*[int(w.shape[1]) for w in ws])
    n_out = (ctypes.c_int * n)(*[int(w.shape[0]) for w in ws])
    lay = MlpTcLayout()
    lib = _capi.load()
    check(lib.smb_mlp_tc_layout_for(n, k_in, n_out, ctypes.byref(lay)), "smb_mlp_tc_layout_for")
    blob = torch.zeros(lay.total_bytes, dtype=torch.uint8)
    fpp = ctypes.POINTER(ctypes.c_float)
    W = (fpp * n)(*[ctypes.cast(w.data_ptr(), fpp) for w in ws])
    B = (fpp * n)(*[ctypes.cast(b.data_ptr(), fpp) for b in bs])
    check(lib.smb_mlp_tc_pack_host(W, B, k_in, n_out, ctypes.byref(lay), ctypes.c_void_p(blob.data_ptr())), "smb_mlp_tc_pack_host")
    return MlpTcPack(blob.to(device), lay, key)


class _ModuleCache:
    """(tag, module) -> pack, weak in the module."""

    def __init__(self) -> None:
        self._d: "weakref.WeakKeyDictionary[torch.nn.Module, Dict[Tuple, MlpTcPack]]" = weakref.WeakKeyDictionary()

    def get(self, tag: Tuple, module: torch.nn.Module) -> Optional[MlpTcPack]:
        return self._d.get(module, {}).get(tag)

    def put(self, tag: Tuple, module: torch.nn.Module, pack: MlpTcPack) -> MlpTcPack:
        self._d.setdefault(module, {})[tag] = pack
        return pack


_mlp_tc_cache = _ModuleCache()


def _param_key(params, device) -> Tuple:
    return tuple((p.data_ptr(), p._version, str(p.device)) for p in params) + (str(device),)


def get_tsr_points_pack(decoder: torch.nn.Module, device: torch.device) -> MlpTcPack:
    """NeRFMLP (network_utils.py:48-79) as a tensor-core MLP for arbitrary positions."""
    ws, bs = decoder_params(decoder)
    key = _param_key((*ws, *bs), device)
    hit = _mlp_tc_cache.get(("tsr",), decoder)
    if hit is None or hit.key != key:
        hit = _mlp_tc_cache.put(("tsr",), decoder, _pack_mlp_tc(ws, bs, device, key))
    return hit


def get_sf3d_points_pack(decoder: torch.nn.Module, device: torch.device) -> MlpTcPack:
    """MaterialMLP heads density + vertex_offset (network.py:158-178) fused into one 3-layer MLP:
    [W0_d;W0_o] (128x120), blockdiag(W1_d, W1_o) (128x128), [[w2_d,0],[0,W2_o]] (4x128)."""
    d = [m for m in decoder.heads["density"] if isinstance(m, torch.nn.Linear)]
    o = [m for m in decoder.heads["vertex_offset"] if isinstance(m, torch.nn.Linear)]
    if len(d) != 3 or len(o) != 3:
        raise NotImplementedError("the CUDA path expects 2 hidden layers per head (config.yaml:50-65)")
    params = [q for m in (*d, *o) for q in (m.weight, m.bias)]
    key = _param_key(params, device)
    hit = _mlp_tc_cache.get(("sf3d",), decoder)
    if hit is None or hit.key != key:
        f = lambda t: t.detach().to("cpu", torch.float32)  # noqa: E731
        w0 = torch.cat([f(d[0].weight), f(o[0].weight)], 0)
        b0 = torch.cat([f(d[0].bias), f(o[0].bias)], 0)
        w1 = torch.block_diag(f(d[1].weight), f(o[1].weight))
        b1 = torch.cat([f(d[1].bias), f(o[1].bias)], 0)
        w2 = torch.zeros(4, 128)
        w2[0, :64] = f(d[2].weight)[0]
        w2[1:, 64:] = f(o[2].weight)
        b2 = torch.cat([f(d[2].bias), f(o[2].bias)], 0)
        hit = _mlp_tc_cache.put(("sf3d",), decoder, _pack_mlp_tc([w0, w1, w2], [b0, b1, b2], device, key))
    return hit


def get_sf3d_head_pack(decoder: torch.nn.Module, name: str, device: torch.device) -> MlpTcPack:
    """One MaterialMLP head with <= 3 outputs (network.py:158-178; e.g. ``features`` / ``perturb_normal``:
    120 -> 64 x n_hidden -> 3) as a tensor-core MLP: its outputs sit in rows 1..3 of the padded last layer (the
    kernel's ``vec`` outputs), row 0 is zero."""
    lin = [m for m in decoder.heads[name] if isinstance(m, torch.nn.Linear)]
    if len(lin) < 2 or lin[0].in_features != 120 or any(m.out_features != 64 for m in lin[:-1]) or lin[-1].out_features > 3:
        raise NotImplementedError(f"head {name!r}: the CUDA path expects 120 -> 64 x n -> (<= 3 outputs), SiLU")
    params = [q for m in lin for q in (m.weight, m.bias)]
    key = _param_key(params, device)
    hit = _mlp_tc_cache.get(("sf3d_head", name), decoder)
    if hit is None or hit.key != key:
        f = lambda t: t.detach().to("cpu", torch.float32)  # noqa: E731
        ws = [f(m.weight) for m in lin[:-1]]
        bs = [f(m.bias) for m in lin[:-1]]
        k = lin[-1].out_features
        w_last, b_last = torch.zeros(4, 64), torch.zeros(4)
        w_last[1 : 1 + k] = f(lin[-1].weight)
        b_last[1 : 1 + k] = f(lin[-1].bias)
        hit = _mlp_tc_cache.put(("sf3d_head", name), decoder, _pack_mlp_tc([*ws, w_last], [*bs, b_last], device, key))
    return hit


def get_sf3d_head_decoder_pack(decoder: torch.nn.Module, name: str, device: torch.device) -> DecoderPack:
    """One MaterialMLP head (network.py:158-178: 120 -> 64 x n_hidden -> k <= 3) in the NeRFMLP blob format of
    ``pack_decoder_host`` (last Linear zero-padded to 4 rows) -- what the lattice tet-grid kernel reads."""
    lin = [m for m in decoder.heads[name] if isinstance(m, torch.nn.Linear)]
    if len(lin) < 3 or lin[0].in_features != 120 or any(m.out_features != 64 for m in lin[:-1]) or lin[-1].out_features > 3:
        raise NotImplementedError(f"head {name!r}: the lattice path expects 120 -> 64 x (n >= 2) -> (<= 3 outputs), SiLU")
    params = [q for m in lin for q in (m.weight, m.bias)]
    key = _param_key(params, device)
    hit = _mlp_tc_cache.get(("sf3d_head_decoder", name), decoder)
    if hit is None or hit.key != key:
        f = lambda t: t.detach().to("cpu", torch.float32)  # noqa: E731
        k = lin[-1].out_features
        w_last, b_last = torch.zeros(4, 64), torch.zeros(4)
        w_last[:k] = f(lin[-1].weight)
        b_last[:k] = f(lin[-1].bias)
        blob, lay = pack_decoder_host([*[f(m.weight) for m in lin[:-1]], w_last], [*[f(m.bias) for m in lin[:-1]], b_last])
        hit = _mlp_tc_cache.put(("sf3d_head_decoder", name), decoder, DecoderPack(blob=blob.to(device), layout=lay, key=key))
    return hit


def query_tetgrid_tc(
    planes: ScenePlanes, packs: Sequence[DecoderPack], n_out: Sequence[int], exp_act: Sequence[bool], out_bias: Sequence[float],
    axis_u: Sequence[torch.Tensor], spatial_dim: Sequence[int], align_corners: bool = True, out_sub: Optional[Sequence[float]] = None,
) -> List[torch.Tensor]:
    """Heads of a MaterialMLP at every vertex of a lattice-ordered grid (``smb_query_tetgrid_tc``): ``axis_u[k]`` is the
    (-1,1) coordinate of lattice index k (slow, mid, fast), ``spatial_dim[k]`` the spatial axis it runs along.
    ``out_sub[h]`` (optional) is subtracted from an exp-activated head after the activation (``density - threshold``).
    Returns one (N, n_out[h]) tensor per head."""
    if planes.planes_cl is None:
        raise ValueError("the lattice tet-grid path reads the fp32 channels-last planes (prepare_planes_cl)")
    dev = planes.planes_cl.device
    nh = len(packs)
    ax = [a.detach().to(device=dev, dtype=torch.float32).contiguous() for a in axis_u]
    ext = [int(a.numel()) for a in ax]
    n = ext[0] * ext[1] * ext[2]
    outs = [torch.empty((n, int(k)), dtype=torch.float32, device=dev) for k in n_out]
    vp, ip, fp = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    blobs = (vp * nh)(*[p.blob.data_ptr() for p in packs])
    lays = (ctypes.POINTER(DecoderLayout) * nh)(*[ctypes.pointer(p.layout) for p in packs])
    with torch.cuda.device(dev):
        check(
            _capi.load().smb_query_tetgrid_tc(
                planes.planes_cl.data_ptr(), planes.Hp, planes.Wp, int(bool(align_corners)), nh, blobs, lays,
                (ip * nh)(*[int(k) for k in n_out]), (ip * nh)(*[int(bool(e)) for e in exp_act]), (fp * nh)(*[float(b) for b in out_bias]),
                (fp * nh)(*[float(x) for x in out_sub]) if out_sub is not None else None,
                (vp * 3)(*[a.data_ptr() for a in ax]), (ip * 3)(*ext), (ip * 3)(*[int(d) for d in spatial_dim]),
                (vp * nh)(*[o.data_ptr() for o in outs]), _stream_ptr(dev),
            ),
            "smb_query_tetgrid_tc",
        )
    return outs


def query_points_tc(
    planes: ScenePlanes, pack: MlpTcPack, positions: torch.Tensor, radius: float, out0_bias: float,
    align_corners: bool, sigmoid_vec: bool, want: Sequence[str] = ("out0_act",),
) -> Dict[str, torch.Tensor]:
    """Tensor-core field query at (n,3) positions.  ``want`` from {out0_raw, out0_act, vec, vec_act}."""
    _require_cuda(positions, "positions")
    pos = positions.detach().to(torch.float32).contiguous().view(-1, 3)
    n, dev = pos.shape[0], pos.device
    widths = {"out0_raw": 1, "out0_act": 1, "vec": 3, "vec_act": 3}
    outs = {k: torch.empty((n, widths[k]), dtype=torch.float32, device=dev) for k in want}
    half = planes.planes_h is not None
    with torch.cuda.device(dev):
        check(
            _capi.load().smb_query_points_tc(
                (planes.planes_h if half else planes.planes_cl).data_ptr(), int(half), planes.Hp, planes.Wp, int(bool(align_corners)), pack.blob.data_ptr(), ctypes.byref(pack.layout),
                float(radius), float(out0_bias), int(bool(sigmoid_vec)), pos.data_ptr(), n, _ptr(outs.get("out0_raw")),
                _ptr(outs.get("out0_act")), _ptr(outs.get("vec")), _ptr(outs.get("vec_act")), _stream_ptr(dev),
            ),
            "smb_query_points_tc",
        )
    return outs
