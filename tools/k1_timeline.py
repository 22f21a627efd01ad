#!/usr/bin/env python3
"""Dev: event timeline of block 0 of the pair kernel (SMB_TC_TRACE=3): who waits for whom."""
import ctypes, os, sys
import numpy as np
import torch
os.environ["SMB_TC_TRACE"] = "3"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import baked_triplane
from sculptmate_b200 import _capi, runtime
from sculptmate_b200.tsr import TSR
R = 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = TSR().to(dev)
pack = runtime.get_decoder_pack(model.decoder, dev)
scene = runtime.prepare_scene(baked_triplane(100).to(dev), pack)
ax = runtime.lattice_axis(R, 0.87, device=dev)
for _ in range(2):
    runtime.query_lattice(scene, pack, ax, R, 0.87, -1.0)
torch.cuda.synchronize()
n = 20 * 512 * 2
buf = (ctypes.c_uint * n)()
lib = _capi.load()
lib.smb_debug_pair_evt.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.smb_debug_pair_evt(buf, n) == 0
a = np.frombuffer(buf, dtype=np.uint32).reshape(20, 512, 2)
g = int(sys.argv[1]) if len(sys.argv) > 1 else 1
# merge events of warpgroup g: consumer warps 4g..4g+3 and issuer 16+g; skip the first pair (start-up)
evs = []
for role in list(range(4 * g, 4 * g + 4)) + [16 + g]:
    for i in range(512):
        ident, clk = int(a[role, i, 0]), int(a[role, i, 1])
        if ident:
            evs.append((clk, role, ident))
evs.sort()
t0 = evs[0][0]
names = {1: "wait lo", 2: "got lo", 3: "rel lo", 4: "wait hi", 5: "got hi", 6: "rel hi", 7: "end"}
lo, hi = int(sys.argv[2]) if len(sys.argv) > 2 else 300, int(sys.argv[3]) if len(sys.argv) > 3 else 520
for clk, role, ident in evs[lo:hi]:
    typ = ident >> 12
    e, s, h = (ident >> 8) & 0xF, (ident >> 4) & 0xF, ident & 0xF
    if typ >= 16:
        print(f"{clk - t0:9d}  ISSUER            MMA e={e} {'XY'[s]} {'lo' if h == 0 else 'hi'}")
    else:
        print(f"{clk - t0:9d}  q{role & 3} {' ' * (12 * (role & 3))}{names[typ]:8s} e={e} {'XY'[s]}")

# ---- per sub-partition: how many of its 4 consumer warps are inside a compute stretch at each instant
print("\nconcurrency of compute stretches per sub-partition (fraction of time with k warps computing):")
for q in range(4):
    iv = []  # (start, end) of compute stretches: rel lo -> wait hi, rel hi -> end
    for gg in range(4):
        role = 4 * gg + q
        st = None
        for i in range(512):
            ident, clk = int(a[role, i, 0]), int(a[role, i, 1])
            if not ident:
                break
            typ = ident >> 12
            if typ in (3, 6):
                st = clk
            elif typ in (4, 7) and st is not None:
                iv.append((st, clk))
                st = None
    if not iv:
        continue
    tmin = max(min(s for s, _ in iv), t0 + 5000)
    tmax = min(max(e for _, e in iv), tmin + 40000)
    ts = np.arange(tmin, tmax, 8)
    k = np.zeros(len(ts), dtype=int)
    for s, e in iv:
        k += (ts >= s) & (ts < e)
    print(f"  q{q}: " + "  ".join(f"k={j}: {100 * np.mean(k == j):4.1f}%" for j in range(5)) + f"   mean {k.mean():.2f}")
