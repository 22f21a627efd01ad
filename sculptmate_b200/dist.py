"""x-slab sharding of extract_mesh across the GPUs of one node (one process per GPU).

The lattice is cut along array axis 0 (the slowest-varying axis = the reference's x,
isosurface.py:33-38).  Rank g owns the cell layers [a_g, b_g) and evaluates density on
planes a_g..b_g inclusive: the one-plane halo is *recomputed* from the broadcast
triplane, never exchanged.  Because the canonical vertex order is plane-major and the
in-plane vertices of plane b_g are the first vertices of rank g+1, a slab numbers them
locally as ``V_g + n``; adding the running offset sum(V_0..V_{g-1}) to every face index
and concatenating slabs in rank order reproduces the single-GPU mesh bit for bit --
no seam search, no welding pass.

Collectives (NCCL over NVLink on the GPU box, gloo in the CPU tests):
  broadcast   triplane (+ decoder parameters when asked)     rank 0 -> all
  all_gather  (V_g, F_g)                                     2 int64 per rank
  send/recv   slab vertices and faces                        ranks -> dst, placed directly
                                                             at their final offsets
There is no collective on the data path of the kernels themselves.
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


def extract_mesh_sharded(
    tsr,
    scene_code: torch.Tensor,
    resolution: int = 256,
    threshold: float = 25.0,
    group: Optional[dist.ProcessGroup] = None,
    dst: int = 0,
    precision: str = "tc",
    broadcast: bool = True,
):
    """TSR.extract_mesh for one scene code with the lattice sharded over the group.
    Returns (v_pos, t_pos_idx) on ``dst`` (device tensors), (None, None) elsewhere."""
    if broadcast and dist.is_available() and dist.is_initialized():
        broadcast_scene(scene_code, None, src=dst, group=group)
    backend = CudaSlabBackend(tsr, scene_code, resolution, threshold, precision)
    verts, faces, _ = gather_slab_meshes(backend, resolution, group=group, dst=dst)
    return verts, faces
