#!/usr/bin/env python3
"""Dev: per-phase clock accumulators of the pair kernel (SMB_TC_TRACE=3 build path)."""
import ctypes, os, sys
import numpy as np
import torch
os.environ["SMB_TC_TRACE"] = "3"
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
scene = runtime.prepare_scene(baked_triplane(100).to(dev), pack)
ax = runtime.lattice_axis(R, 0.87, device=dev)
for _ in range(2):
    runtime.query_lattice(scene, pack, ax, R, 0.87, -1.0)
torch.cuda.synchronize()
n = 148 * 16 * 16 + 148 * 16
buf = (ctypes.c_uint * n)()
lib = _capi.load()
lib.smb_debug_pair_prof.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.smb_debug_pair_prof(buf, n) == 0
allv = np.frombuffer(buf, dtype=np.uint32)
a = allv[: 148 * 16 * 16].reshape(148, 16, 16).astype(np.float64)
mm = allv[148 * 16 * 16 :].reshape(148, 4, 4).astype(np.float64)
print(f'issuer: MMA issue -> commit observed, mean clocks: lo {2 * mm[:, :, 0].sum() / mm[:, :, 2].sum():.0f}  hi {2 * mm[:, :, 1].sum() / mm[:, :, 2].sum():.0f}  (commits {mm[:, :, 2].sum():.0f})')
names = ["total", "L0 wait t_full", "L0 compute+st+arrive", "H wait acc lo", "H ld lo + wait::ld", "H compute lo", "H wait acc hi", "H ld hi/st lo/wait::ld",
         "H compute hi + st", "H wait::st + arrive", "head wait lo", "head rest (incl wait hi)", "L0 geometry"]
tot = a[:, :, 0].mean()
print(f"mean total clocks per consumer warp: {tot:.0f}")
for i, nm in enumerate(names):
    if i:
        print(f"  {nm:28s} {a[:, :, i].mean():12.0f}  {100 * a[:, :, i].mean() / tot:5.1f} %")
print("per warpgroup (mean over CTAs and the 4 warps): total, wait acc lo, wait acc hi, compute lo+hi")
for g in range(4):
    w = a[:, 4 * g : 4 * g + 4, :]
    print(f"  wg{g}: total {w[:, :, 0].mean():9.0f} (min {w[:, :, 0].min():9.0f} max {w[:, :, 0].max():9.0f})  wait lo {w[:, :, 3].mean():8.0f}  wait hi {w[:, :, 6].mean():8.0f}  compute {w[:, :, 5].mean() + w[:, :, 8].mean():9.0f}")
print("per sub-partition q (mean over CTAs and warpgroups): total, wait lo, wait hi, compute")
for q in range(4):
    w = a[:, q::4, :]
    print(f"  q{q}: total {w[:, :, 0].mean():9.0f}  wait lo {w[:, :, 3].mean():8.0f}  wait hi {w[:, :, 6].mean():8.0f}  compute {w[:, :, 5].mean() + w[:, :, 8].mean():9.0f}")
