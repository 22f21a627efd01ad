#!/usr/bin/env python3
"""Dev timing of the lattice kernel alone (CUDA events, L2 flushed between runs):
    python tools/bench_lattice.py [R] [iters]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import baked_triplane  # noqa: E402
from sculptmate_b200 import runtime  # noqa: E402
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
out = torch.empty((R, R, R), dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(iters + 3):
    flush.fill_(i & 0xFF)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    runtime.query_lattice(scene, pack, axis, R, 0.87, -1.0, out=out)
    b.record()
    torch.cuda.synchronize()
    if i >= 3:
        ts.append(a.elapsed_time(b))
ms = float(np.median(ts))
print(f"lattice_tc R={R}: median {ms:.3f} ms (min {min(ts):.3f})  {R**3 / ms / 1e6:.2f} Gpts/s  {81408 * R**3 / ms / 1e9:.1f} TFLOP/s  checksum {float(out.double().sum()):.6e}")
