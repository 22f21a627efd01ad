#!/usr/bin/env python3
"""Multi-GPU parity check (one process per GPU, NCCL): the x-slab sharded extract_mesh must
reproduce the single-GPU mesh bit for bit.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_sharded.py [R]
Rank 0 prints one line per check and exits non-zero on a mismatch."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import baked_triplane  # noqa: E402
from sculptmate_b200.dist import broadcast_scene, extract_mesh_sharded  # noqa: E402
from sculptmate_b200.tsr import TSR  # noqa: E402


def main() -> int:
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = TSR().to(dev)
    ok = True
    for seed in (100, 101):
        # only rank 0 holds the real scene: the broadcast is part of what is checked
        tp = baked_triplane(seed).to(dev) if rank == 0 else torch.zeros(3, 40, 64, 64, device=dev)
        broadcast_scene(tp, model.decoder, src=0)
        thr = float(model.renderer.query_lattice(model.decoder, tp, 64).median())
        for transport in ("p2p", "nccl", "p2p"):
            v, f = extract_mesh_sharded(model, tp, R, thr, broadcast=False, transport=transport)
            if rank == 0:
                v1, f1 = model.extract_mesh_tensors(tp, R, thr)
                same = v.shape == v1.shape and f.shape == f1.shape and torch.equal(v, v1) and torch.equal(f, f1)
                print(f"sharded x{world} R={R} seed={seed} {transport}: V={v.shape[0]} F={f.shape[0]} bit-exact vs 1 GPU: {same}", flush=True)
                ok &= bool(same)
    # a bigger surface than the mapped buffers hold: the p2p transport must regrow collectively
    tp = baked_triplane(100).to(dev) if rank == 0 else torch.zeros(3, 40, 64, 64, device=dev)
    broadcast_scene(tp, None, src=0)
    for Rb in (R, R + 32):
        thr = float(model.renderer.query_lattice(model.decoder, tp, 64).median())
        v, f = extract_mesh_sharded(model, tp, Rb, thr, broadcast=False)
        if rank == 0:
            v1, f1 = model.extract_mesh_tensors(tp, Rb, thr)
            same = torch.equal(v, v1) and torch.equal(f, f1)
            print(f"sharded x{world} R={Rb} regrow p2p: V={v.shape[0]} F={f.shape[0]} bit-exact vs 1 GPU: {same}", flush=True)
            ok &= bool(same)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
