#!/usr/bin/env python3
"""Dev timing of the host-to-host C-ABI call alone (smb_extract_mesh_host: pinned triplane in, mesh in pinned host memory
out), for sweeping the slab-pipeline knobs in fresh processes:
    SMB_PIPE_SLABS=3 SMB_PIPE_RATIO=0.45 python tools/bench_e2e.py [R] [iters]"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import RADIUS, baked_triplane, decoder_numpy  # noqa: E402
from sculptmate_b200 import _capi  # noqa: E402
from sculptmate_b200.tsr import TSR  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
lib = _capi.load()
_, ws, bs = decoder_numpy(0)
fpp = ctypes.POINTER(ctypes.c_float)
W = (fpp * 10)(*[w.ctypes.data_as(fpp) for w in ws])
B = (fpp * 10)(*[b.ctypes.data_as(fpp) for b in bs])
ex = ctypes.c_void_p()
_capi.check(lib.smb_extractor_create(W, B, 9, RADIUS, -1.0, 64, 64, ctypes.byref(ex)), "create")
seeds = [100, 101, 102, 103]
host_tp = [np.ascontiguousarray(baked_triplane(s).numpy()) for s in seeds]
torch.manual_seed(0)
model = TSR().to(dev)
thr = [float(model.renderer.query_lattice(model.decoder, torch.from_numpy(t).to(dev), 64).median()) for t in host_tp]
vp, fp_ = fpp(), ctypes.POINTER(ctypes.c_int64)()
nv, nt = ctypes.c_int64(), ctypes.c_int64()
pin = fpp()
_capi.check(lib.smb_extractor_pinned_input(ex, ctypes.byref(pin)), "pinned_input")
pin_np = np.ctypeslib.as_array(pin, shape=host_tp[0].shape)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i32 in (0, 1):
    _capi.check(lib.smb_extractor_set_faces_i32(ex, i32), "faces_i32")
    ts = []
    for i in range(iters + 5):
        np.copyto(pin_np, host_tp[i % 4])
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = lib.smb_extract_mesh_host(ex, pin, R, thr[i % 4], ctypes.byref(vp), ctypes.byref(fp_), ctypes.byref(nv), ctypes.byref(nt))
        dt = time.perf_counter() - t0
        _capi.check(rc, "extract")
        if i >= 5:
            ts.append(dt * 1e3)
    print(f"e2e R={R} faces={'int32' if i32 else 'int64'} slabs={os.environ.get('SMB_PIPE_SLABS', 'default')} ratio={os.environ.get('SMB_PIPE_RATIO', 'default')}: "
          f"mean {np.mean(ts):.3f} ms  p50 {np.median(ts):.3f}  min {min(ts):.3f}  max {max(ts):.3f}  V={nv.value} F={nt.value}")
lib.smb_extractor_destroy(ex)
