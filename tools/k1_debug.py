#!/usr/bin/env python3
"""Dev: run the tensor-core lattice kernel on a series of shapes, synchronising and timing each launch."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import baked_triplane
from sculptmate_b200 import runtime
from sculptmate_b200.tsr import TSR

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = TSR().to(dev)
pack = runtime.get_decoder_pack(model.decoder, dev)
scene = runtime.prepare_scene(baked_triplane(100).to(dev), pack)
cases = [(8, 0, 8), (8, 2, 4), (8, 0, 8), (8, 0, 1), (32, 0, 32), (32, 3, 7), (64, 0, 64), (100, 0, 100), (256, 0, 16), (256, 0, 256)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for R, x0, nx in cases:
    ax = runtime.lattice_axis(R, 0.87, device=dev)
    ref = runtime.query_lattice(scene, pack, ax, R, 0.87, -1.0, x_begin=x0, nx=nx, precision="fp32")
    torch.cuda.synchronize()
    t = time.time()
    try:
        out = runtime.query_lattice(scene, pack, ax, R, 0.87, -1.0, x_begin=x0, nx=nx, precision="tc")
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        import ctypes
        from sculptmate_b200 import _capi
        _capi.load().smb_debug_pair_dump()
        print(f"R={R} x0={x0} nx={nx}: FAILED after {time.time() - t:.2f} s: {str(e).splitlines()[0]}", flush=True)
        break
    print(f"R={R} x0={x0} nx={nx}: ok {1e3 * (time.time() - t):.2f} ms  max rel err {float((out / ref - 1).abs().max()):.2e}", flush=True)
