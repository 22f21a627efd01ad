#!/usr/bin/env python3
"""CPU study for the next K1 design: how much precision would fp16 accumulators (tcgen05 kind::f16 with
D = f16, 64 TMEM columns per tile instead of 96) cost?  Emulates the kernel's arithmetic in numpy -- fp16 weights (W/2
folded), fp16 activations, SiLU as h + h*tanh(h), four K=16 MMA steps per layer -- once with an fp32 accumulator
(today) and once with an accumulator rounded to fp16 after every K=16 step, against the fp32 reference MLP.

    python tools/study_fp16_accumulate.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sculptmate_b200.tsr import NeRFMLP  # noqa: E402


def silu_half_arg(h):
    return h + h * np.tanh(h)


def run(ws, bs, x, acc16: bool):
    # layer 0 in fp32 from exact features (the kernel interpolates projected planes in fp32)
    h = (x @ (0.5 * ws[0]).T + 0.5 * bs[0]).astype(np.float32)
    a = silu_half_arg(h).astype(np.float16)
    for l in range(1, len(ws)):
        last = l == len(ws) - 1
        W = ((1.0 if last else 0.5) * ws[l]).astype(np.float16).astype(np.float32)
        A = a.astype(np.float32)
        acc = np.zeros((A.shape[0], W.shape[0]), np.float32)
        for k in range(0, 64, 16):
            acc = acc + A[:, k : k + 16] @ W[:, k : k + 16].T
            if acc16:
                acc = acc.astype(np.float16).astype(np.float32)
        if last:
            return acc[:, 0] + bs[l][0]
        a = silu_half_arg(acc + 0.5 * bs[l]).astype(np.float16)


def reference(ws, bs, x):
    h = x.astype(np.float64)
    for l in range(len(ws)):
        h = h @ ws[l].astype(np.float64).T + bs[l]
        if l != len(ws) - 1:
            h = h / (1 + np.exp(-h))
    return h[:, 0]


for seed, wscale in ((0, 1.0), (1, 1.0), (2, 2.0), (3, 3.0)):
    torch.manual_seed(seed)
    dec = NeRFMLP(dict(in_channels=120, n_neurons=64, n_hidden_layers=9, activation="silu"))
    sd = dec.state_dict()
    ws = [sd[f"layers.{i}.weight"].numpy() * (wscale if 0 < i < 18 else 1.0) for i in range(0, 20, 2)]
    bs = [sd[f"layers.{i}.bias"].numpy() for i in range(0, 20, 2)]
    x = np.random.RandomState(seed).randn(20000, 120).astype(np.float32) * 0.5
    ref = reference(ws, bs, x)
    e32 = np.abs(run(ws, bs, x, False) - ref)
    e16 = np.abs(run(ws, bs, x, True) - ref)
    print(f"seed {seed} hidden-weight scale {wscale}: logit range [{ref.min():.3f}, {ref.max():.3f}]  "
          f"fp32 acc: max {e32.max():.2e} rms {np.sqrt((e32**2).mean()):.2e}   fp16 acc: max {e16.max():.2e} rms {np.sqrt((e16**2).mean()):.2e}  "
          f"(x{e16.max() / e32.max():.1f})")
