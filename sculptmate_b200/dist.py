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

``p2p`` (default on CUDA, world > 1) -- fused with the emit kernel, NO collective on the data path.  The
  destination rank owns the mesh buffers and every rank a small control block (cudaMalloc + CUDA IPC, mapped into
  every process once).  Per call each rank runs  lattice (+ sign ballot) -> mc_count -> ``peer_publish_counts``
  (its counts + a sequence flag stored over NVLink into every rank's control block) -> ``mc_emit`` in gather mode:
  the kernel waits ON THE DEVICE for the lower ranks' flags, sums their counts to get its output offsets and stores
  vertices and int32 faces (global ids) straight into the destination's buffers over NVLink peer memory ->
  ``peer_signal_done`` (a flag in the destination's block).  The destination's ``peer_wait_all`` kernel waits for all
  flags, releases the peers for the next call and sums the counts.  No staging buffer, no send/recv pass, no NCCL
  call, no host round trip between the kernels; one host sync per call (the sizes).  Round 1 used an NCCL
  all_gather for the counts and an all_reduce as completion fence (two collective launches + their latency per
  call) and carried int64 faces on the wire.
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
    """Rank ``src`` (rank inside ``group``) -> all: the triplane (1.97 MB fp32) and, optionally, the decoder
    parameters (0.17 MB).  In place."""
    gsrc = _global_rank(group, src)  # torch.distributed.broadcast takes a GLOBAL rank; `src` here is group-relative
    dist.broadcast(scene_code, src=gsrc, group=group)
    if decoder is not None:
        for p in decoder.parameters():
            dist.broadcast(p.data, src=gsrc, group=group)


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
    """Persistent state of the ``p2p`` transport for one (process group, destination):
      * the destination's double-buffered mesh storage, mapped into every other process (CUDA IPC);
      * one control block per rank (counts, sequence flags), every block mapped by every process, and the device
        table of those pointers;
      * device / pinned totals.
    Capacities grow by re-running ``setup_mesh`` collectively (every rank reads the same totals each call)."""

    SETS = 2  # results stay valid until the call after the next one
    CTRL_WORDS = 128

    def __init__(self, device: torch.device, group=None, dst: int = 0):
        from . import _capi

        self.device, self.group, self.dst = device, group, dst
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise ValueError("the peer-flag gather supports up to 16 ranks per node")
        self.vcap = self.fcap = 0
        self.allocs: list = []  # dst: [(_DevAlloc verts, _DevAlloc faces)] * SETS
        self.ptrs: list = []  # every rank: [(verts ptr, faces ptr)] * SETS in THIS process' address space
        self.turn = 0
        self.seq = 0
        self.wide: list = []  # dst: per set, the int64 copy of the faces (when the caller wants the reference's LongTensor)
        self.counts_dev = torch.zeros(4, dtype=torch.int64, device=device)
        self.totals_dev = torch.zeros(4, dtype=torch.int64, device=device)
        self.totals_pin = torch.zeros(4, dtype=torch.int64).pin_memory()
        self._lib = _capi.load()
        self._setup_ctrl()

    def _setup_ctrl(self) -> None:
        """Collective, once: allocate this rank's control block, map everybody else's."""
        import ctypes

        from . import _capi

        lib = self._lib
        with torch.cuda.device(self.device):
            self.ctrl = _DevAlloc(8 * self.CTRL_WORDS)
            torch.as_tensor(_CaiView(self.ctrl, self.ctrl.ptr, (self.CTRL_WORDS,), "<i8"), device=self.device).zero_()
            torch.cuda.synchronize(self.device)
            mine = self.ctrl.handle().to(self.device)
            allh = torch.zeros(64 * self.world, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allh, mine, group=self.group)
            hb = allh.cpu().numpy().tobytes()
            self.ctrl_ptrs = []
            for g in range(self.world):
                if g == self.rank:
                    self.ctrl_ptrs.append(self.ctrl.ptr)
                    continue
                h = (ctypes.c_ubyte * 64).from_buffer_copy(hb[64 * g : 64 * (g + 1)])
                q = ctypes.c_void_p()
                _capi.check(lib.smb_ipc_open(h, ctypes.byref(q)), "smb_ipc_open")
                self.ctrl_ptrs.append(int(q.value))
            self.ctrl_table = torch.tensor(self.ctrl_ptrs, dtype=torch.int64, device=self.device)
        dist.barrier(self.group)  # every block is zeroed and mapped before anyone publishes into it

    def setup_mesh(self, vcap: int, fcap: int) -> None:
        """Collective.  (Re)allocate the destination buffers and map them everywhere."""
        import ctypes

        from . import _capi

        lib = self._lib
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
                    av, af = _DevAlloc(12 * vcap), _DevAlloc(24 * fcap)  # faces sized for int64; int32 uses half
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
                        q = ctypes.c_void_p()
                        _capi.check(lib.smb_ipc_open(h, ctypes.byref(q)), "smb_ipc_open")
                        out.append(int(q.value))
                    self.ptrs.append(tuple(out))
        self.vcap, self.fcap = int(vcap), int(fcap)
        self.wide = []
        dist.barrier(self.group)

    def views(self, k: int, V: int, F: int, faces_dtype: torch.dtype):
        av, af = self.allocs[k]
        verts = torch.as_tensor(_CaiView(av, av.ptr, (V, 3), "<f4"), device=self.device)
        faces = torch.as_tensor(_CaiView(af, af.ptr, (F, 3), "<i4" if faces_dtype == torch.int32 else "<i8"), device=self.device)
        return verts, faces


def _extract_mesh_p2p(tsr, scene_code, resolution, threshold, group, dst, precision, faces_dtype=torch.int64, wire_i32=True,
                      broadcast=False, phases=None):
    """One sharded extract with the peer-flag gather.  ``wire_i32``: faces cross NVLink as int32 (global ids < 2^31) and
    are widened on the destination when ``faces_dtype`` is int64.  ``phases`` (optional dict) receives CUDA events
    at the phase boundaries of this rank (developer timing)."""
    from . import _capi, runtime
    from ._capi import MC_AFFINE, MC_COALESCE, MC_DIV, MC_FACES_I32, MC_FLIP

    dev = scene_code.device
    key = (id(group), dst, str(dev))
    cache = tsr.__dict__.setdefault("_peer_gather", {})
    pg = cache.get(key)
    if pg is None:
        pg = cache[key] = PeerGather(dev, group, dst)
    world, rank = pg.world, pg.rank
    a, b = slab_partition(resolution, world)[rank]
    nx, R, last = b - a + 1, resolution, rank == world - 1
    lib = pg._lib
    tsr.set_marching_cubes_resolution(R)
    r = tsr.renderer.cfg.radius
    i32 = bool(wire_i32) or faces_dtype == torch.int32
    # ranks other than the destination store into PEER memory: coalesced 128-byte stores (staged in shared memory)
    flags = MC_FLIP | MC_DIV | MC_AFFINE | (MC_FACES_I32 if i32 else 0) | (MC_COALESCE if rank != dst else 0)

    def mark(name):
        if phases is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            phases[name] = e

    with torch.cuda.device(dev):
        st = runtime._stream_ptr(dev)
        mark("start")
        pg.seq += 1
        _capi.check(lib.smb_peer_wait_release(pg.ctrl.ptr, pg.seq - 1, st), "smb_peer_wait_release")
        if broadcast:
            broadcast_scene(scene_code, None, src=dst, group=group)
            if rank != dst:
                tsr.renderer._scene_ref = None  # the tensor was rewritten in place by NCCL: never reuse planes prepared from its old content
        mark("scene")
        with torch.no_grad():
            # the tensor-core lattice kernel also ballots the marching-cubes sign masks of the slab into the workspace
            fused = precision == "tc"
            slab = tsr.renderer.query_lattice(tsr.decoder, scene_code, R, axis_u=tsr._axis(R, dev), x_begin=a, nx=nx, precision=precision,
                                              mc_signs=(float(threshold), 1.0) if fused else None)
        mark("lattice")
        w = runtime._mc_cache.get(dev, (nx, R, R))
        w.generation += 1
        if fused and w.signed_matches(slab, threshold, 1.0):
            _capi.check(lib.smb_mc_count_presigned(nx, R, R, int(last), w.ws.data_ptr(), w.ws.numel(), pg.counts_dev.data_ptr(), st), "smb_mc_count_presigned")
        else:
            _capi.check(lib.smb_mc_count(slab.data_ptr(), nx, R, R, float(threshold), 1.0, int(last), w.ws.data_ptr(), w.ws.numel(),
                                         pg.counts_dev.data_ptr(), st), "smb_mc_count")
        first = True
        while True:
            if not first:  # capacity grown (collectively): a fresh sequence number, only publish + emit are repeated
                pg.seq += 1
                _capi.check(lib.smb_peer_wait_release(pg.ctrl.ptr, pg.seq - 1, st), "smb_peer_wait_release")
            _capi.check(lib.smb_peer_publish_counts(pg.counts_dev.data_ptr(), pg.ctrl_table.data_ptr(), rank, world, pg.seq, st), "smb_peer_publish_counts")
            mark("count")
            if pg.vcap > 0:
                k = pg.turn % pg.SETS
                pv, pf = pg.ptrs[k]
                _capi.check(
                    lib.smb_mc_emit_gather_flags(slab.data_ptr(), nx, R, R, float(threshold), 1.0, a, int(last), flags, float(R - 1.0),
                                                 float(r - (-r)), float(-r), w.ws.data_ptr(), pg.ctrl.ptr, pg.seq, rank,
                                                 pv, pg.vcap, pf, pg.fcap, st),
                    "smb_mc_emit_gather_flags",
                )
            mark("emit")
            _capi.check(lib.smb_peer_signal_done(pg.ctrl_ptrs[dst], rank, pg.seq, st), "smb_peer_signal_done")
            _capi.check(lib.smb_peer_wait_all(pg.ctrl.ptr, pg.ctrl_table.data_ptr(), world, pg.seq, int(rank == dst), pg.totals_dev.data_ptr(), st),
                        "smb_peer_wait_all")
            mark("gathered")
            pg.totals_pin.copy_(pg.totals_dev, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            V, F, err = int(pg.totals_pin[0]), int(pg.totals_pin[1]), int(pg.totals_pin[2])
            if err:
                raise RuntimeError("sharded extract_mesh: a peer did not arrive within the device-side timeout")
            if i32 and V >= 2**31:
                raise ValueError("vertex ids do not fit int32; call with wire_i32=False")
            if V <= pg.vcap and F <= pg.fcap and pg.vcap > 0:
                break
            pg.setup_mesh(max(pg.vcap, V * 5 // 4 + 4096), max(pg.fcap, F * 5 // 4 + 4096))  # every rank sees the same totals
            first = False
        pg.turn += 1
        if V == 0 or F == 0:
            # same exception types as the single-GPU path (skimage's, isosurface.py:46-48), decided on the GLOBAL
            # value range so that every rank raises the same one
            lo, hi = runtime.grid_minmax(slab, float(threshold), 1.0)
            mm = torch.tensor([-lo, hi], dtype=torch.float32, device=dev)
            dist.all_reduce(mm, op=dist.ReduceOp.MAX, group=group)
            lo, hi = -float(mm[0]), float(mm[1])
            if lo > 0.0 or hi < 0.0:
                raise ValueError("Surface level must be within volume data range.")
            raise RuntimeError("No surface found at the given iso value.")
        if rank == dst:
            verts, faces = pg.views(k, V, F, torch.int32 if i32 else torch.int64)
            if i32 and faces_dtype == torch.int64:
                # widened once on the destination (the wire carried 12 B per triangle) into a persistent buffer of the same
                # double-buffered lifetime as the mapped ones
                wide = pg.wide[k] if len(pg.wide) > k and pg.wide[k] is not None and pg.wide[k].shape[0] >= F else None
                if wide is None:
                    while len(pg.wide) <= k:
                        pg.wide.append(None)
                    wide = pg.wide[k] = torch.empty((pg.fcap, 3), dtype=torch.int64, device=dev)
                _capi.check(lib.smb_mesh_faces_i64(faces.data_ptr(), F, wide.data_ptr(), st), "smb_mesh_faces_i64")
                faces = wide[:F]
            mark("end")
            return verts, faces
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
    faces_dtype: torch.dtype = torch.int64,
    phases: Optional[dict] = None,
):
    """TSR.extract_mesh for one scene code with the lattice sharded over the group.
    Returns (v_pos, t_pos_idx) on ``dst`` (device tensors), (None, None) elsewhere.  With the
    ``p2p`` transport the vertex tensor (and int32 faces) are views of persistent buffers, valid until the
    call after the next one.  Single process / world size 1: the single-GPU path (``extract_mesh_tensors``)."""
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if not multi and scene_code.is_cuda:
        return tsr.extract_mesh_tensors(scene_code, resolution, threshold, precision=precision, faces_dtype=faces_dtype)
    if multi and transport == "p2p" and scene_code.is_cuda:
        return _extract_mesh_p2p(tsr, scene_code, resolution, threshold, group, dst, precision, faces_dtype=faces_dtype,
                                 broadcast=broadcast, phases=phases)
    if broadcast and multi:
        broadcast_scene(scene_code, None, src=dst, group=group)
    backend = CudaSlabBackend(tsr, scene_code, resolution, threshold, precision)
    verts, faces, _ = gather_slab_meshes(backend, resolution, group=group, dst=dst)
    return verts, faces
