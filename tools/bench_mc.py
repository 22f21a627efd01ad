#!/usr/bin/env python3
"""Dev timing of the marching-cubes kernels alone on the bench scene (CUDA events, L2 flushed between runs):
    python tools/bench_mc.py [R] [iters]
Prints count(+totals) and emit separately for int64 and int32 faces, plus a checksum of the mesh so that variants
(SMB_MC_BATCH=8|16|32 in a fresh process) can be compared for equality."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import baked_triplane  # noqa: E402
from sculptmate_b200 import _capi, runtime  # noqa: E402
from sculptmate_b200.tsr import TSR  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = TSR().to(dev)
pack = runtime.get_decoder_pack(model.decoder, dev)
model.set_marching_cubes_resolution(R)
axis = model._axis(R, dev)
scene = runtime.prepare_scene(baked_triplane(100).to(dev), pack, want_cl=False, want_q=True)
grid = torch.empty((R, R, R), dtype=torch.float32, device=dev)
runtime.query_lattice(scene, pack, axis, R, 0.87, -1.0, out=grid)
thr = float(grid.flatten()[:: max(1, R**3 // (1 << 22))].median())  # bench.py's rule: the median density
pend = runtime.mc_count(grid, sub=thr, sign=1.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

# work distribution: items (crossing samples + active cells) per batch of W consecutive words, from the word records
nw = R * R * ((R + 31) // 32)
rec_off = ((nw + 1) * 4 + 255) // 256 * 256
rec = pend.wsp.ws[rec_off: rec_off + nw * 32].view(torch.int32).view(nw, 8).cpu().numpy().astype(np.uint32)
popc = np.vectorize(lambda x: bin(int(x)).count("1"), otypes=[np.int64])
lut = np.array([bin(i).count("1") for i in range(65536)], dtype=np.int64)
pc = lambda a: lut[a & 0xFFFF] + lut[a >> 16]
ns = pc(rec[:, 4] | rec[:, 5] | rec[:, 6])
nc = pc(rec[:, 7])
for W in (32, 16, 8):
    nb = nw // W
    s_ = ns[: nb * W].reshape(nb, W).sum(1)
    c_ = nc[: nb * W].reshape(nb, W).sum(1)
    it = (s_ + 31) // 32 + (c_ + 31) // 32
    print(f"W={W}: batches {nb}, non-empty {(it > 0).mean():.2f}, group iterations total {it.sum()} (ideal {(ns.sum() + 31) // 32 + (nc.sum() + 31) // 32}), "
          f"per batch mean {it.mean():.2f} max {it.max()}, items per batch max {int((s_ + c_).max())}")


def ev():
    return torch.cuda.Event(enable_timing=True)


for flags, name in ((0, "int64"), (_capi.MC_FACES_I32, "int32")):
    fdt = torch.int32 if flags else torch.int64
    verts = torch.empty((pend.nverts, 3), dtype=torch.float32, device=dev)
    faces = torch.empty((pend.ntris, 3), dtype=fdt, device=dev)
    tc, te = [], []
    for i in range(iters + 3):
        flush.fill_(i & 0xFF)
        a, b, c = ev(), ev(), ev()
        a.record()
        lib = _capi.load()
        w = pend.wsp
        runtime._launch_count(lib, grid, R, R, R, thr, 1.0, True, w, runtime._stream_ptr(dev), False)
        b.record()
        runtime.check(
            lib.smb_mc_emit_bounded(grid.data_ptr(), R, R, R, thr, 1.0, 0, 1, flags, 1.0, 1.0, 0.0, 0, w.ws.data_ptr(), verts.data_ptr(),
                                    pend.nverts, faces.data_ptr(), pend.ntris, runtime._stream_ptr(dev)),
            "emit",
        )
        c.record()
        torch.cuda.synchronize()
        if i >= 3:
            tc.append(a.elapsed_time(b))
            te.append(b.elapsed_time(c))
    cs = (float(verts.double().sum()), int(faces.long().sum()), int((faces.long() * torch.arange(1, 4, device=dev)).sum() % (1 << 61)))
    print(f"mc R={R} {name}: signs+count+totals {np.median(tc)*1e3:.1f} us  emit {np.median(te)*1e3:.1f} us (min {min(te)*1e3:.1f})  "
          f"V={pend.nverts} F={pend.ntris}  checksum {cs[0]:.6e} {cs[1]} {cs[2]}")
