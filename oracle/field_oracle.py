"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, fp32) of the reference's
triplane field query.  Never imported by the product package ``sculptmate_b200``;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may use it, and only as the checker / the baseline.

Pinned: yes -- ``tests/test_oracle_golden.py`` checks every function here against
fixtures produced by the *unmodified* reference (``oracle/make_golden.py`` imports
it from /root/reference through ``oracle/ref_shim.py``), and
``tests/test_oracle_vs_reference.py`` re-checks live when /root/reference exists.

What each function restates (paths relative to /root/reference):

  grid_axis / grid_vertices   TripoSR/tsr/models/isosurface.py:25-39
                              (torch.linspace + meshgrid(indexing="ij"), x slowest)
  scale_tensor                TripoSR/tsr/utils.py:222-231
  grid_sample_bilinear        aten grid_sampler_2d, bilinear, zeros padding,
                              align_corners False (TripoSR/tsr/models/nerf_renderer.py:61-66)
                              or True (StableFast/sf3d/system.py:190-195)
  triplane_features           TripoSR/tsr/models/nerf_renderer.py:56-68
                              (plane p samples (x,y),(x,z),(y,z); concat plane-major)
  nerf_mlp                    TripoSR/tsr/models/network_utils.py:49-79,116-124
  query_triplane              TripoSR/tsr/models/nerf_renderer.py:41-91
  rays_intersect_bbox         TripoSR/tsr/utils.py:115-149
  render_rays                 TripoSR/tsr/models/nerf_renderer.py:93-152 (sample positions, alpha compositing)
  sf3d_query_triplane         StableFast/sf3d/system.py:170-198
  material_mlp_head           StableFast/sf3d/models/network.py:158-178,191-208
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

F32 = np.float32


# ---------------------------------------------------------------- lattice ---
def grid_axis(resolution: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    """torch.linspace(lo, hi, R) in fp32 (isosurface.py:30-32).

    aten computes ``step = (hi-lo)/(R-1)`` in fp32 and fills the first half as
    ``lo + step*i`` and the second half as ``hi - step*(R-1-i)``, each as ONE fused
    multiply-add (single rounding): emulated by forming the product in float64 (exact for
    two fp32 factors) and rounding the sum once.  Checked against torch.linspace for
    every R in 2..400 (tests/test_oracle_vs_reference.py).
    """
    R = int(resolution)
    if R == 1:
        return np.array([lo], dtype=F32)
    lo32, hi32 = F32(lo), F32(hi)
    step = F32((hi32 - lo32) / F32(R - 1))
    i = np.arange(R)
    first = (np.float64(lo32) + np.float64(step) * i).astype(F32)
    second = (np.float64(hi32) - np.float64(step) * (R - 1 - i)).astype(F32)
    return np.where(i < R // 2, first, second).astype(F32)


def grid_vertices(resolution: int) -> np.ndarray:
    """(R^3, 3) fp32 lattice, row (i*R+j)*R+k = (x_i, y_j, z_k) (isosurface.py:33-38)."""
    a = grid_axis(resolution)
    x, y, z = np.meshgrid(a, a, a, indexing="ij")
    return np.stack([x.reshape(-1), y.reshape(-1), z.reshape(-1)], axis=-1).astype(F32)


def scale_tensor(dat: np.ndarray, inp_scale: Tuple[float, float], tgt_scale: Tuple[float, float]) -> np.ndarray:
    """utils.py:222-231 with python-float scales applied as fp32 scalars (aten CPU)."""
    if inp_scale is None:
        inp_scale = (0, 1)
    if tgt_scale is None:
        tgt_scale = (0, 1)
    dat = dat.astype(F32)
    dat = (dat - F32(inp_scale[0])) / F32(inp_scale[1] - inp_scale[0])
    dat = dat * F32(tgt_scale[1] - tgt_scale[0]) + F32(tgt_scale[0])
    return dat.astype(F32)


# ------------------------------------------------------------ grid_sample ---
def _unnormalize(coord: np.ndarray, size: int, align_corners: bool) -> np.ndarray:
    coord = coord.astype(F32)
    if align_corners:
        # ((coord + 1) / 2) * (size - 1)
        return ((coord + F32(1)) * F32((size - 1) / 2.0)).astype(F32)
    # aten CPU: (coord + 1) * (size / 2) - 0.5   [== ((coord+1)*size - 1)/2]
    return ((coord + F32(1)) * F32(size / 2.0) - F32(0.5)).astype(F32)


def grid_sample_bilinear(plane: np.ndarray, u: np.ndarray, v: np.ndarray, align_corners: bool) -> np.ndarray:
    """plane (C,H,W) fp32; u -> W axis, v -> H axis, both (N,) in [-1,1]. Returns (N,C)."""
    C, H, W = plane.shape
    ix = _unnormalize(u, W, align_corners)
    iy = _unnormalize(v, H, align_corners)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    wx1 = (ix - x0).astype(F32)
    wx0 = (F32(1) - wx1).astype(F32)
    wy1 = (iy - y0).astype(F32)
    wy0 = (F32(1) - wy1).astype(F32)
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    out = np.zeros((u.shape[0], C), dtype=F32)
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xs = x0 + dx
            ys = y0 + dy
            ok = (xs >= 0) & (xs < W) & (ys >= 0) & (ys < H)
            xs_c = np.clip(xs, 0, W - 1)
            ys_c = np.clip(ys, 0, H - 1)
            vals = plane[:, ys_c, xs_c].T  # (N,C)
            w = (wy * wx).astype(F32) * ok.astype(F32)
            out += vals * w[:, None]
    return out.astype(F32)


_PLANE_AXES = ((0, 1), (0, 2), (1, 2))  # nerf_renderer.py:57-60 / sf3d/system.py:186-189


def triplane_features(triplane: np.ndarray, pos_normalized: np.ndarray, align_corners: bool = False) -> np.ndarray:
    """triplane (3,Cp,Hp,Wp); pos_normalized (N,3) in [-1,1] -> (N, 3*Cp), plane-major."""
    feats = []
    for p, (a, b) in enumerate(_PLANE_AXES):
        feats.append(grid_sample_bilinear(triplane[p], pos_normalized[:, a], pos_normalized[:, b], align_corners))
    return np.concatenate(feats, axis=1).astype(F32)


# -------------------------------------------------------------------- MLP ---
def silu(x: np.ndarray) -> np.ndarray:
    x = x.astype(F32)
    return (x / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


def nerf_mlp(x: np.ndarray, weights: Sequence[np.ndarray], biases: Sequence[np.ndarray]) -> Dict[str, np.ndarray]:
    """network_utils.py:116-124: Linear/SiLU chain, last Linear has no activation.

    ``weights[l]`` is (out,in) as nn.Linear stores it; y = x @ W.T + b.
    """
    h = x.astype(F32)
    n = len(weights)
    for l, (W, b) in enumerate(zip(weights, biases)):
        h = (h @ W.astype(F32).T + b.astype(F32)).astype(F32)
        if l != n - 1:
            h = silu(h)
    return {"density": h[:, 0:1], "features": h[:, 1:4]}


def query_triplane(
    positions: np.ndarray,
    triplane: np.ndarray,
    weights: Sequence[np.ndarray],
    biases: Sequence[np.ndarray],
    radius: float = 0.87,
    density_bias: float = -1.0,
    chunk: int = 1 << 16,
) -> Dict[str, np.ndarray]:
    """nerf_renderer.py:41-91 (feature_reduction=concat, exp / sigmoid activations)."""
    in_shape = positions.shape[:-1]
    pos = positions.reshape(-1, 3).astype(F32)
    pos = scale_tensor(pos, (-radius, radius), (-1, 1))
    dens: List[np.ndarray] = []
    feat: List[np.ndarray] = []
    for s in range(0, max(1, pos.shape[0]), chunk):
        f = triplane_features(triplane.astype(F32), pos[s : s + chunk], align_corners=False)
        o = nerf_mlp(f, weights, biases)
        dens.append(o["density"])
        feat.append(o["features"])
    density = np.concatenate(dens, axis=0)
    features = np.concatenate(feat, axis=0)
    out = {
        "density": density,
        "features": features,
        "density_act": np.exp(density + F32(density_bias), dtype=F32),
        "color": (F32(1) / (F32(1) + np.exp(-features, dtype=F32))).astype(F32),
    }
    return {k: v.reshape(*in_shape, -1).astype(F32) for k, v in out.items()}


def grid_density(
    resolution: int,
    triplane: np.ndarray,
    weights: Sequence[np.ndarray],
    biases: Sequence[np.ndarray],
    radius: float = 0.87,
) -> np.ndarray:
    """The density half of TSR.extract_mesh (system.py:171-184): (R,R,R) density_act."""
    pos = scale_tensor(grid_vertices(resolution), (0, 1), (-radius, radius))
    out = query_triplane(pos, triplane, weights, biases, radius=radius)
    return out["density_act"].reshape(resolution, resolution, resolution)


# ------------------------------------------------------------------- SF3D ---
def rays_intersect_bbox(rays_o: np.ndarray, rays_d: np.ndarray, radius: float, near: float = 0.0, valid_thresh: float = 0.01):
    """utils.py:115-149 in fp32: slab test against the box tightened by (1 - 1e-3)."""
    o = rays_o.reshape(-1, 3).astype(np.float32)
    d = rays_d.reshape(-1, 3).astype(np.float32)
    d = np.where(np.abs(d) < np.float32(1e-6), np.float32(1e-6), d)
    hi = np.float32(np.float32(1.0 - 1.0e-3) * np.float32(radius))
    lo = np.float32(np.float32(1.0 - 1.0e-3) * np.float32(-radius))
    a = ((hi - o) / d).astype(np.float32)
    b = ((lo - o) / d).astype(np.float32)
    t_near = np.maximum(np.minimum(a, b).max(axis=-1), np.float32(near))
    t_far = np.maximum(a, b).min(axis=-1)
    valid = (t_far - t_near) > np.float32(valid_thresh)
    t_near = np.where(valid, t_near, np.float32(0)).astype(np.float32)
    t_far = np.where(valid, t_far, np.float32(0)).astype(np.float32)
    return t_near, t_far, valid


def render_rays(triplane: np.ndarray, rays_o: np.ndarray, rays_d: np.ndarray, weights, biases, radius: float = 0.87,
                density_bias: float = -1.0, num_samples: int = 128) -> np.ndarray:
    """nerf_renderer.py:93-152 for rays that all hit the box (the only case the reference supports)."""
    shape = rays_o.shape[:-1]
    o = rays_o.reshape(-1, 3).astype(np.float32)
    d = rays_d.reshape(-1, 3).astype(np.float32)
    t_near, t_far, valid = rays_intersect_bbox(o, d, radius)
    assert valid.all(), "the reference's _forward needs every ray to hit the box"
    t_vals = np.linspace(0.0, 1.0, num_samples + 1).astype(np.float32)
    t_mid = ((t_vals[:-1] + t_vals[1:]) / np.float32(2.0)).astype(np.float32)
    z = (t_near[:, None] * (np.float32(1) - t_mid[None]) + t_far[:, None] * t_mid[None]).astype(np.float32)
    xyz = (o[:, None, :] + z[..., None] * d[:, None, :]).astype(np.float32)
    out = query_triplane(xyz.reshape(-1, 3), triplane, weights, biases, radius=radius, density_bias=density_bias)
    sigma = out["density_act"].reshape(-1, num_samples)
    color = out["color"].reshape(-1, num_samples, 3)
    deltas = (t_vals[1:] - t_vals[:-1]).astype(np.float32)
    alpha = (np.float32(1) - np.exp(-deltas[None] * sigma)).astype(np.float32)
    trans = np.cumprod((np.float32(1) - alpha[:, :-1] + np.float32(1e-10)).astype(np.float32), axis=-1, dtype=np.float32)
    accum = np.concatenate([np.ones_like(alpha[:, :1]), trans], axis=-1)
    w = (alpha * accum).astype(np.float32)
    rgb = (w[..., None] * color).sum(axis=-2, dtype=np.float32)
    opacity = w.sum(axis=-1, dtype=np.float32)
    return (rgb + (np.float32(1) - opacity)[:, None]).astype(np.float32).reshape(*shape, 3)


def sf3d_query_triplane(positions: np.ndarray, triplane: np.ndarray, radius: float = 0.87) -> np.ndarray:
    """sf3d/system.py:170-198 for one un-batched triplane: (N,3) -> (N, 3*Cp)."""
    pos = scale_tensor(positions.reshape(-1, 3), (-radius, radius), (-1, 1))
    return triplane_features(triplane.astype(F32), pos, align_corners=True)


def material_mlp_head(
    x: np.ndarray, weights: Sequence[np.ndarray], biases: Sequence[np.ndarray], out_bias: float = 0.0, activation: str | None = None
) -> np.ndarray:
    """One head of MaterialMLP (sf3d/models/network.py:158-178,200-207)."""
    h = x.astype(F32)
    n = len(weights)
    for l, (W, b) in enumerate(zip(weights, biases)):
        h = (h @ W.astype(F32).T + b.astype(F32)).astype(F32)
        if l != n - 1:
            h = silu(h)
    h = (h + F32(out_bias)).astype(F32)
    if activation in (None, "none"):
        return h
    if activation in ("trunc_exp", "exp"):
        return np.exp(h, dtype=F32)
    if activation == "sigmoid":
        return (F32(1) / (F32(1) + np.exp(-h, dtype=F32))).astype(F32)
    raise ValueError(activation)


# ------------------------------------------------------- synthetic inputs ---
def decoder_params_from_state_dict(sd) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """layers.{0,2,...,18}.{weight,bias} -> ([W0..W9], [b0..b9]) as numpy fp32."""
    ws, bs = [], []
    for i in range(0, 20, 2):
        ws.append(np.asarray(sd[f"layers.{i}.weight"], dtype=F32))
        bs.append(np.asarray(sd[f"layers.{i}.bias"], dtype=F32))
    return ws, bs
