#!/usr/bin/env python3
"""Developer tool: per-step clock64 timeline of block 0 / warpgroup 0 of the lattice kernel
(SMB_TC_TRACE=1 kernel variant).  python tools/trace_lattice.py [R]"""
import ctypes, os, sys
import numpy as np
import torch
os.environ["SMB_TC_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import baked_triplane
from sculptmate_b200 import _capi, runtime
from sculptmate_b200.tsr import TSR
R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = TSR().to(dev)
pack = runtime.get_decoder_pack(model.decoder, dev)
model.set_marching_cubes_resolution(R)
axis = model._axis(R, dev)
scene = runtime.prepare_scene(baked_triplane(100).to(dev), pack, want_cl=False, want_q=True)
for _ in range(3):
    runtime.query_lattice(scene, pack, axis, R, 0.87, -1.0)
torch.cuda.synchronize()
lib = _capi.load()
n = 4 * 512 * 4
buf = (ctypes.c_longlong * n)()
lib.smb_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.smb_debug_read_trace(buf, n) == 0
t = np.array(buf[:], dtype=np.int64).reshape(4, 512, 4)
base = t[:, 40, 0].min()
print("ev  l s | per warp q: wait_start, wait_done-wait_start, compute, publish   (cycles; relative start)")
for ev in range(40, 100):
    l, s = (ev // 2) % 10, ev % 2
    row = []
    for q in range(4):
        a, b, c, d = t[q, ev]
        if l == 0:
            b = a
        row.append(f"{a-base:7d} w{b-a:5d} c{c-b:5d} p{d-c:4d}")
    print(f"{ev:3d} {l:2d} {s} | " + " | ".join(row))
