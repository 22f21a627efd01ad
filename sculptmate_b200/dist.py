"""x-slab sharding of extract_mesh across the GPUs of one node (one process per GPU).

The lattice is cut along array axis 0 (the slowest-varying axis = the reference's x,
isosurface.py:33-38).  Rank g owns the cell layers [a_g, b_g) and evaluates density on
planes a_g..b_g inclusive: the one-plane halo is *recomputed* from the broadcast
triplane, never exchanged.  Because the canonical vertex order is plane-major and the
in-plane vertices of plane b_g are the first vertices of rank g+1, a slab numbers them
locally as ``V_g + n``; adding the running offset sum(V_0..V_{g-1}) to every face index
and concatenating slabs in rank order reproduces the single-GPU mesh bit for bit --
no seam search, no welding pass.

Two transports for the gather:

``p2p`` (default on CUDA, world > 1) -- fused with the emit kernel.  The destination rank owns
  the mesh buffers (cudaMalloc + CUDA IPC, mapped into every process once); per call each rank
  runs  lattice -> mc_count -> [NCCL all_gather of the 4-int64 counts, device to device] ->
  mc_emit in gather mode, whose kernel derives its output offsets from the gathered counts ON THE
  DEVICE and stores vertices and faces (with global ids) straight into the destination's buffers
  over NVLink peer memory -> [tiny NCCL all_reduce as the completion fence].  No staging buffer,
  no send/recv pass, no host round trip between the kernels; one host sync per call (the sizes).
``nccl`` -- count -> all_gather -> emit locally -> grouped isend/irecv of the slab meshes to
  their final offsets.  Also what the gloo CPU tests exercise with the oracle as the backend.

Other collectives: ``broadcast`` of the triplane (+ decoder parameters when asked), rank 0 -> all.
"""
from __future__ import annotations

from typing import List, Optional, Protocol, Tuple

import torch
import torch.distributed as dist


def slab_partition(resolution: int, world_size: int) -> List[Tuple[int, int]]:
    """Cell-layer ranges [a_g, b_g) over the R-1 layers, as even as possible, contiguous."""
    cells = resolution - 1
    if world_size < 1 or cells < world_size:
        raise ValueError(f"cannot split {cells} cell layers over {world_size} ranks")
    base, extra = divmod(cells, world_size)
    out, a = [], 0
    for g in range(world_size):
        b = a + base + (1 if g < extra else 0)
        out.append((a, b))
        a = b
    return out


class SlabBackend(Protocol):
    """What a rank must be able to do for its slab.  The product backend is CUDA
    (`CudaSlabBackend`); the CPU tests plug in the oracle to exercise the protocol."""

    device: torch.device

    def count(self, x_begin: int, nx: int, emit_last_plane: bool) -> Tuple[int, int]: ...

    def emit(self, x_origin: int, vertex_id_offset: int, verts_out: torch.Tensor, faces_out: torch.Tensor) -> None: ...


class CudaSlabBackend:
    def __init__(self, tsr, scene_code: torch.Tensor, resolution: int, threshold: float, precision: str = "tc"):
        from . import runtime  # noqa: F401  (fails loudly if the CUDA library is missing)

        self.tsr, self.scene_code, self.R, self.threshold, self.precision = tsr, scene_code, resolution, threshold, precision
        self.device = scene_code.device
        self._pend = None

    def count(self, x_begin: int, nx: int, emit_last_plane: bool) -> Tuple[int, int]:
        from . import runtime

        t = self.tsr
        t.set_marching_cubes_resolution(self.R)
        with torch.no_grad():
            slab = t.renderer.query_lattice(
                t.decoder, self.scene_code, self.R, axis_u=t._axis(self.R, self.device), x_begin=x_begin, nx=nx,
                precision=self.precision,
            )
        self._pend = runtime.mc_count(slab, sub=float(self.threshold), sign=1.0, emit_last_plane=emit_last_plane)
        return self._pend.nverts, self._pend.ntris

    def emit(self, x_origin: int, vertex_id_offset: int, verts_out: torch.Tensor, faces_out: torch.Tensor) -> None:
        from . import runtime
        from ._capi import MC_AFFINE, MC_DIV, MC_FLIP

        r = self.tsr.renderer.cfg.radius
        runtime.mc_emit(
            self._pend, x_origin=x_origin, flags=MC_FLIP | MC_DIV | MC_AFFINE, vdiv=float(self.R - 1.0),
            vmul=float(r - (-r)), vadd=float(-r), vertex_id_offset=vertex_id_offset,
            verts_out=verts_out, faces_out=faces_out,
        )


def gather_slab_meshes(
    backend: SlabBackend,
    resolution: int,
    group: Optional[dist.ProcessGroup] = None,
    dst: int = 0,
) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], List[Tuple[int, int]]]:
    """Run count -> all_gather -> emit -> gather for this rank's slab.

    Returns (verts, faces, counts) with the merged mesh on ``dst`` (None elsewhere);
    ``counts`` is [(V_g, F_g)] for every rank.
    """
    solo = not (dist.is_available() and dist.is_initialized())  # single process: one slab, no collective
    world = 1 if solo else dist.get_world_size(group)
    rank = 0 if solo else dist.get_rank(group)
    dev = backend.device
    parts = slab_partition(resolution, world)
    a, b = parts[rank]
    last = rank == world - 1
    V, F = backend.count(a, b - a + 1, last)

    if solo:
        counts = [(int(V), int(F))]
    else:
        mine = torch.tensor([V, F], dtype=torch.int64, device=dev)
        allc = torch.empty(2 * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, mine, group=group)
        allc = allc.cpu().view(world, 2)
        counts = [(int(v), int(f)) for v, f in allc.tolist()]
    v_off = [0]
    f_off = [0]
    for v, f in counts:
        v_off.append(v_off[-1] + v)
        f_off.append(f_off[-1] + f)

    if rank == dst:
        verts = torch.empty((v_off[-1], 3), dtype=torch.float32, device=dev)
        faces = torch.empty((f_off[-1], 3), dtype=torch.int64, device=dev)
        my_v = verts[v_off[rank] : v_off[rank + 1]]
        my_f = faces[f_off[rank] : f_off[rank + 1]]
    else:
        verts = faces = None
        my_v = torch.empty((V, 3), dtype=torch.float32, device=dev)
        my_f = torch.empty((F, 3), dtype=torch.int64, device=dev)
    # rank dst writes its slab straight into the merged buffers
    backend.emit(a, v_off[rank], my_v, my_f)

    ops = []
    if rank == dst:
        for g in range(world):
            if g == dst:
                continue
            if counts[g][0]:
                ops.append(dist.P2POp(dist.irecv, verts[v_off[g] : v_off[g + 1]], _global_rank(group, g), group))
            if counts[g][1]:
                ops.append(dist.P2POp(dist.irecv, faces[f_off[g] : f_off[g + 1]], _global_rank(group, g), group))
    else:
        if V:
            ops.append(dist.P2POp(dist.isend, my_v, _global_rank(group, dst), group))
        if F:
            ops.append(dist.P2POp(dist.isend, my_f, _global_rank(group, dst), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return verts, faces, counts


def _global_rank(group: Optional[dist.ProcessGroup], group_rank: int) -> int:
    return group_rank if group is None else dist.get_global_rank(group, group_rank)


def broadcast_scene(scene_code: torch.Tensor, decoder: Optional[torch.nn.Module] = None, src: int = 0, group=None) -> None:
    """Rank ``src`` -> all: the triplane (1.97 MB fp32) and, optionally, the decoder
    parameters (0.17 MB).  In place."""
    dist.broadcast(scene_code, src=src, group=group)
    if decoder is not None:
        for p in decoder.parameters():
            dist.broadcast(p.data, src=src, group=group)


# ------------------------------------------------------------------ p2p transport
class _CaiView:
    """Exposes a region of a raw device allocation to torch (zero-copy) through
    ``__cuda_array_interface__``; keeps the owning allocation alive."""

    def __init__(self, owner, ptr: int, shape, typestr: str):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class _DevAlloc:
    def __init__(self, nbytes: int):
        import ctypes

        from . import _capi

        self._lib = _capi.load()
        p = ctypes.c_void_p()
        _capi.check(self._lib.smb_dev_alloc(int(nbytes), ctypes.byref(p)), "smb_dev_alloc")
        self.ptr, self.nbytes = int(p.value), int(nbytes)

    def handle(self) -> torch.Tensor:
        import ctypes

        from . import _capi

        h = (ctypes.c_ubyte * 64)()
        _capi.check(self._lib.smb_ipc_export(self.ptr, h), "smb_ipc_export")
        return torch.tensor(list(h), dtype=torch.uint8)

    def __del__(self):
        try:
            self._lib.smb_dev_free(self.ptr)
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


class PeerGather:
    """Persistent state of the ``p2p`` transport for one (process group, destination): the
    destination's double-buffered mesh storage, its mapping in every other process, and the
    device/pinned count buffers.  Capacities grow by re-running ``setup`` collectively."""

    SETS = 2  # results stay valid until the call after the next one

    def __init__(self, device: torch.device, group=None, dst: int = 0):
        self.device, self.group, self.dst = device, group, dst
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.vcap = self.fcap = 0
        self.allocs: list = []  # dst: [(_DevAlloc verts, _DevAlloc faces)] * SETS
        self.ptrs: list = []  # every rank: [(verts ptr, faces ptr)] * SETS in THIS process' address space
        self.turn = 0
        self.counts_dev = torch.zeros(4, dtype=torch.int64, device=device)
        self.all_counts_dev = torch.zeros(4 * self.world, dtype=torch.int64, device=device)
        self.all_counts_pin = torch.zeros(4 * self.world, dtype=torch.int64).pin_memory()
        self.token = torch.zeros(1, dtype=torch.int32, device=device)

    def setup(self, vcap: int, fcap: int) -> None:
        """Collective.  (Re)allocate the destination buffers and map them everywhere."""
        import ctypes

        from . import _capi

        lib = _capi.load()
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)  # nobody is still writing into the old buffers
        if self.rank != self.dst:
            for pv, pf in self.ptrs:
                lib.smb_ipc_close(pv)
                lib.smb_ipc_close(pf)
        self.ptrs, self.allocs = [], []
        handles = torch.zeros(self.SETS * 2 * 64, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            if self.rank == self.dst:
                hs = []
                for _ in range(self.SETS):
                    av, af = _DevAlloc(12 * vcap), _DevAlloc(24 * fcap)
                    self.allocs.append((av, af))
                    self.ptrs.append((av.ptr, af.ptr))
                    hs += [av.handle(), af.handle()]
                handles.copy_(torch.cat(hs))
            dist.broadcast(handles, src=_global_rank(self.group, self.dst), group=self.group)
            if self.rank != self.dst:
                hb = handles.cpu().numpy().tobytes()
                for k in range(self.SETS):
                    out = []
                    for j in range(2):
                        h = (ctypes.c_ubyte * 64).from_buffer_copy(hb[(2 * k + j) * 64 : (2 * k + j + 1) * 64])
                        p = ctypes.c_void_p()
                        _capi.check(lib.smb_ipc_open(h, ctypes.byref(p)), "smb_ipc_open")
                        out.append(int(p.value))
                    self.ptrs.append(tuple(out))
        self.vcap, self.fcap = int(vcap), int(fcap)
        dist.barrier(self.group)

    def views(self, k: int, V: int, F: int):
        av, af = self.allocs[k]
        verts = torch.as_tensor(_CaiView(av, av.ptr, (V, 3), "<f4"), device=self.device)
        faces = torch.as_tensor(_CaiView(af, af.ptr, (F, 3), "<i8"), device=self.device)
        return verts, faces


def _extract_mesh_p2p(tsr, scene_code, resolution, threshold, group, dst, precision):
    import ctypes  # noqa: F401

    from . import _capi, runtime
    from ._capi import MC_AFFINE, MC_DIV, MC_FLIP

    dev = scene_code.device
    key = (id(group), dst, str(dev))
    cache = tsr.__dict__.setdefault("_peer_gather", {})
    pg = cache.get(key)
    if pg is None:
        pg = cache[key] = PeerGather(dev, group, dst)
    world, rank = pg.world, pg.rank
    a, b = slab_partition(resolution, world)[rank]
    nx, R, last = b - a + 1, resolution, rank == world - 1
    lib = _capi.load()
    tsr.set_marching_cubes_resolution(R)
    with torch.no_grad():
        # the tensor-core lattice kernel also ballots the marching-cubes sign masks of the slab into the workspace
        fused = precision == "tc"
        slab = tsr.renderer.query_lattice(tsr.decoder, scene_code, R, axis_u=tsr._axis(R, dev), x_begin=a, nx=nx, precision=precision,
                                          mc_signs=(float(threshold), 1.0) if fused else None)
    ws, _, _ = runtime._mc_cache.get(dev, (nx, R, R))
    r = tsr.renderer.cfg.radius
    flags = MC_FLIP | MC_DIV | MC_AFFINE

    def counts_to_host():
        pg.all_counts_pin.copy_(pg.all_counts_dev, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        c = pg.all_counts_pin.view(world, 4)
        return int(c[:, 0].sum()), int(c[:, 1].sum())

    with torch.cuda.device(dev):
        st = runtime._stream_ptr(dev)
        if fused:
            _capi.check(lib.smb_mc_count_presigned(nx, R, R, int(last), ws.data_ptr(), ws.numel(), pg.counts_dev.data_ptr(), st), "smb_mc_count_presigned")
        else:
            _capi.check(lib.smb_mc_count(slab.data_ptr(), nx, R, R, float(threshold), 1.0, int(last), ws.data_ptr(), ws.numel(),
                                         pg.counts_dev.data_ptr(), st), "smb_mc_count")
        dist.all_gather_into_tensor(pg.all_counts_dev, pg.counts_dev, group=group)
        totals = None
        if pg.vcap == 0:  # first call: learn the sizes, then map buffers with 25 % head room
            totals = counts_to_host()
            pg.setup(totals[0] * 5 // 4 + 4096, totals[1] * 5 // 4 + 4096)
        while True:
            k = pg.turn % pg.SETS
            pv, pf = pg.ptrs[k]
            _capi.check(
                lib.smb_mc_emit_gather(slab.data_ptr(), nx, R, R, float(threshold), 1.0, a, int(last), flags, float(R - 1.0),
                                       float(r - (-r)), float(-r), ws.data_ptr(), pg.all_counts_dev.data_ptr(), rank,
                                       pv, pg.vcap, pf, pg.fcap, st),
                "smb_mc_emit_gather",
            )
            dist.all_reduce(pg.token, group=group)  # completion fence: every slab has been stored
            if totals is None:
                totals = counts_to_host()
            if totals[0] <= pg.vcap and totals[1] <= pg.fcap:
                break
            pg.setup(totals[0] * 5 // 4 + 4096, totals[1] * 5 // 4 + 4096)  # outgrown (every rank sees the same counts)
        pg.turn += 1
        if rank == dst:
            torch.cuda.current_stream(dev).synchronize()
            return pg.views(k, totals[0], totals[1])
    return None, None


def extract_mesh_sharded(
    tsr,
    scene_code: torch.Tensor,
    resolution: int = 256,
    threshold: float = 25.0,
    group: Optional[dist.ProcessGroup] = None,
    dst: int = 0,
    precision: str = "tc",
    broadcast: bool = True,
    transport: str = "p2p",
):
    """TSR.extract_mesh for one scene code with the lattice sharded over the group.
    Returns (v_pos, t_pos_idx) on ``dst`` (device tensors), (None, None) elsewhere.  With the
    ``p2p`` transport the tensors are views of persistent buffers, valid until the call after
    the next one."""
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if broadcast and multi:
        broadcast_scene(scene_code, None, src=dst, group=group)
    if multi and transport == "p2p" and scene_code.is_cuda:
        return _extract_mesh_p2p(tsr, scene_code, resolution, threshold, group, dst, precision)
    backend = CudaSlabBackend(tsr, scene_code, resolution, threshold, precision)
    verts, faces, _ = gather_slab_meshes(backend, resolution, group=group, dst=dst)
    return verts, faces
