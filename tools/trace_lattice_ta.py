#!/usr/bin/env python3
"""Developer tool: per-layer-step clock64 timeline of block 0 of the default lattice kernel (SMB_TC_TRACE=2 build of
lattice_tc_ta_kernel): for every consumer warp, when a step starts waiting for its accumulator, when the accumulator is
ready, when the activations are stored, when the next MMA has been issued.  Prints per-warp phase durations and, per
SM sub-partition, how many of its consumer warps are inside an activation stretch over time.

    python tools/trace_lattice_ta.py [R] [wgs]"""
import ctypes
import os
import sys

import numpy as np
import torch

os.environ["SMB_TC_TRACE"] = "2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import baked_triplane  # noqa: E402
from sculptmate_b200 import _capi, runtime  # noqa: E402
from sculptmate_b200.tsr import TSR  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
WG = int(sys.argv[2]) if len(sys.argv) > 2 else 5
os.environ["SMB_TC_TA_WG"] = str(WG)
STEPS = 160
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
n = 5 * 4 * STEPS * 4
buf = (ctypes.c_longlong * n)()
lib.smb_debug_read_trace_ta.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.smb_debug_read_trace_ta(buf, n) == 0
t = np.array(buf[:], dtype=np.int64).reshape(5, 4, STEPS, 4)[:WG]
t0 = t[:, :, 0, 0].min()
t = t - t0
lo, hi = 40, 140  # steady state: skip the first 4 tiles
wait = t[:, :, lo:hi, 1] - t[:, :, lo:hi, 0]
comp = t[:, :, lo:hi, 2] - t[:, :, lo:hi, 1]
sync = t[:, :, lo:hi, 3] - t[:, :, lo:hi, 2]
gap = t[:, :, lo + 1 : hi + 1, 0] - t[:, :, lo:hi, 3]
step = t[:, :, lo + 1 : hi + 1, 0] - t[:, :, lo:hi, 0]
kind = np.arange(lo, hi) % 10  # 0 = layer 0 (table), 1..8 hidden epilogues, 9 = head
print(f"R={R}, {WG} warpgroups, block 0, steps {lo}..{hi} (cycles)")
print("phase medians over all consumer warps:  wait-for-accumulator  activation-stretch  st.wait+barrier+MMA-issue  to-next-step  whole-step")
for name, sel in (("hidden epilogue", (kind >= 1) & (kind <= 8)), ("layer 0", kind == 0), ("head", kind == 9)):
    f = lambda a: f"{np.median(a[:, :, sel]):8.0f} (p90 {np.percentile(a[:, :, sel], 90):6.0f})"  # noqa: E731
    print(f"  {name:16s} {f(wait)} {f(comp)} {f(sync)} {f(gap)} {f(step)}")
tile = t[:, :, lo + 10 : hi + 1 : 10, 0] - t[:, :, lo : hi - 9 : 10, 0]
print(f"cycles per tile (10 steps): median {np.median(tile):.0f}  -> SFU-only floor {9 * 64 * 8 * WG} for {WG} warps per sub-partition")
# concurrency: for each sub-partition q, the number of warps inside an activation stretch, sampled over the window
a, b = t[:, :, lo, 0].max(), t[:, :, hi - 1, 3].min()
ts = np.linspace(a, b, 4000)
print("sub-partition: share of time with k consumer warps inside an activation stretch (k = 0..%d), mean k" % WG)
for q in range(4):
    inside = np.zeros_like(ts)
    for g in range(WG):
        s0, s1 = t[g, q, :, 1], t[g, q, :, 2]
        inside += ((ts[:, None] >= s0[None]) & (ts[:, None] < s1[None])).sum(1)
    hist = [float((inside == k).mean()) for k in range(WG + 1)]
    print(f"  q={q}: " + " ".join(f"{h:5.2f}" for h in hist) + f"   mean {inside.mean():.2f}")
print("first steady-state steps of sub-partition 0 (start, +wait, +compute, +sync per warpgroup):")
for sidx in range(lo, lo + 12):
    print(f"  step {sidx:3d} k={sidx % 10}: " + " | ".join(f"{t[g, 0, sidx, 0]:7d} w{t[g, 0, sidx, 1] - t[g, 0, sidx, 0]:5d} c{t[g, 0, sidx, 2] - t[g, 0, sidx, 1]:5d} s{t[g, 0, sidx, 3] - t[g, 0, sidx, 2]:4d}" for g in range(WG)))
