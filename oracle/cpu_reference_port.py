"""TEST/BENCH INFRASTRUCTURE ONLY -- the reference's CPU extract_mesh path as a PORT,
for timing beside the GPU numbers on boxes where /root/reference does not exist.

The reference is pure Python over aten (no native sources to compile), so its CPU path
is fully described by the aten calls it makes.  This module issues the same calls in the
same order and chunking -- multi-threaded aten, fp32:
    F.grid_sample x3 planes (one batched call) -> rearrange -> nn.Sequential(Linear, SiLU...)
    per 8192-point chunk (nerf_renderer.py:56-79, generate.py:11), torch.cat
    (utils.py:205), exp / sigmoid (nerf_renderer.py:82-87),
followed by the in-repo CPU marching cubes (oracle/mc_oracle.c; scikit-image, which the
reference calls at isosurface.py:46-48, is not installed -- labelled "oracle MC, not
skimage" wherever it is reported).  tests/test_oracle_vs_reference.py checks this port
against the live reference when it is available.
"""
from __future__ import annotations

import time
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import mc_oracle

RADIUS = 0.87


def _scale(dat, inp, tgt):
    dat = (dat - inp[0]) / (inp[1] - inp[0])
    return dat * (tgt[1] - tgt[0]) + tgt[0]


def make_layers(ws: List[np.ndarray], bs: List[np.ndarray]) -> torch.nn.Sequential:
    mods = []
    for l, (w, b) in enumerate(zip(ws, bs)):
        lin = torch.nn.Linear(w.shape[1], w.shape[0])
        with torch.no_grad():
            lin.weight.copy_(torch.as_tensor(w))
            lin.bias.copy_(torch.as_tensor(b))
        mods.append(lin)
        if l != len(ws) - 1:
            mods.append(torch.nn.SiLU(inplace=True))
    return torch.nn.Sequential(*mods)


@torch.no_grad()
def query_triplane(layers, positions: torch.Tensor, triplane: torch.Tensor, chunk: int = 8192) -> Dict[str, torch.Tensor]:
    pos = _scale(positions.view(-1, 3), (-RADIUS, RADIUS), (-1, 1))
    dens, feat = [], []
    for s in range(0, max(1, pos.shape[0]), chunk):
        x = pos[s : s + chunk]
        idx = torch.stack((x[..., [0, 1]], x[..., [0, 2]], x[..., [1, 2]]), dim=-3)
        out = F.grid_sample(triplane, idx[:, None], align_corners=False, mode="bilinear")  # (3,Cp,1,N)
        out = out[:, :, 0].permute(2, 0, 1).reshape(x.shape[0], -1)  # N (Np Cp)
        y = layers(out)
        dens.append(y[..., 0:1])
        feat.append(y[..., 1:4])
    density = torch.cat(dens, 0)
    features = torch.cat(feat, 0)
    return {
        "density": density,
        "features": features,
        "density_act": torch.exp(density + (-1.0)),
        "color": torch.sigmoid(features),
    }


@torch.no_grad()
def extract_mesh_slab(
    layers, triplane: torch.Tensor, resolution: int, threshold: float, x_begin: int, nx: int
) -> Tuple[np.ndarray, np.ndarray, Dict[str, float]]:
    """The reference extract_mesh body (system.py:171-189) restricted to nx x-planes of the
    lattice (a bounded sample of the full R^3 job).  Returns verts, faces, timings."""
    R = resolution
    ax = torch.linspace(0, 1, R)
    t0 = time.perf_counter()
    x, y, z = torch.meshgrid(ax[x_begin : x_begin + nx], ax, ax, indexing="ij")
    verts = torch.cat([x.reshape(-1, 1), y.reshape(-1, 1), z.reshape(-1, 1)], dim=-1)
    dens = query_triplane(layers, _scale(verts, (0, 1), (-RADIUS, RADIUS)), triplane)["density_act"]
    t1 = time.perf_counter()
    level = -(-(dens - threshold)).view(nx, R, R)
    v, f, c = mc_oracle.marching_cubes_slab(level.numpy(), x_origin=x_begin, flags=mc_oracle.FLIP | mc_oracle.DIV, vdiv=float(R - 1.0))
    v = _scale(torch.from_numpy(v), (0, 1), (-RADIUS, RADIUS)).numpy()
    t2 = time.perf_counter()
    return v, f, {"query_s": t1 - t0, "mc_s": t2 - t1, "points": float(nx) * R * R}
