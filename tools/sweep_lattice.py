#!/usr/bin/env python3
"""Dev sweep of the lattice kernel's build-time variants (SMB_TC_TA_BIAS x SMB_TC_TA_POLY) in ONE process:
time at R (CUDA events, L2 flushed) and error of density_act / raw logit against the fp32 CUDA-core kernel.
    python tools/sweep_lattice.py [R] [iters] [weight_scale]
weight_scale > 1 multiplies the hidden weights so that pre-activations leave the |h| < 1 comfort zone of a
random-init decoder and exercise the whole range of the tanh polynomial."""
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
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
wscale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = TSR().to(dev)
if wscale != 1.0:
    with torch.no_grad():
        for i in range(0, 18, 2):
            model.decoder.layers[i].weight.mul_(wscale)
            model.decoder.layers[i].bias.mul_(4.0)
pack = runtime.get_decoder_pack(model.decoder, dev)
model.set_marching_cubes_resolution(R)
axis = model._axis(R, dev)
scene = runtime.prepare_scene(baked_triplane(100).to(dev), pack, want_cl=True, want_q=True)
Rr = min(R, 128)
axis_r = model._axis(Rr, dev)
ref_act, ref_raw = runtime.query_lattice(scene, pack, axis_r, Rr, 0.87, -1.0, precision="fp32", want_raw=True)
out = torch.empty((R, R, R), dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(f"R={R} iters={iters} wscale={wscale}; fp32 logits: min {float(ref_raw.min()):.3f} max {float(ref_raw.max()):.3f} std {float(ref_raw.std()):.4f}")
# (bias_mma, poly, stagger_clk, xu_tokens)
configs = [(0, 0, 0, 0), (1, 0, 0, 0), (0, 4, 0, 0), (1, 4, 0, 0), (0, 0, 650, 0), (0, 0, 0, 1), (0, 0, 0, 2), (0, 0, 0, 3), (1, 4, 0, 2)]
if os.environ.get("SWEEP_CONFIGS"):
    configs = [tuple(int(x) for x in c.split(",")) for c in os.environ["SWEEP_CONFIGS"].split(";")]
for cfg in configs:
    bias, poly, stag, tok = cfg[:4]
    pipe = cfg[4] if len(cfg) > 4 else 0
    os.environ["SMB_TC_TA_PIPE"] = str(pipe)
    os.environ["SMB_TC_TA_BIAS"] = str(bias)
    os.environ["SMB_TC_TA_POLY"] = str(poly)
    os.environ["SMB_TC_TA_STAGGER"] = str(stag)
    os.environ["SMB_TC_TA_TOKENS"] = str(tok)
    act, raw = runtime.query_lattice(scene, pack, axis_r, Rr, 0.87, -1.0, want_raw=True)
    torch.cuda.synchronize()
    e_raw = float((raw - ref_raw).abs().max())
    e_rms = float((raw - ref_raw).pow(2).mean().sqrt())
    e_rel = float(((act - ref_act).abs() / ref_act.abs()).max())
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
    print(f"bias_mma={bias} poly={poly}/16 stagger={stag} tokens={tok} pipe={pipe}: {ms:.3f} ms (min {min(ts):.3f}) {81408 * R**3 / ms / 1e9:.1f} TFLOP/s | logit err max {e_raw:.2e} rms {e_rms:.2e} | density_act max rel {e_rel:.2e}", flush=True)
