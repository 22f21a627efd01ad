"""Volume renderer: TriplaneNeRFRenderer.forward/_forward and rays_intersect_bbox against the
fixture produced by the unmodified reference (oracle/make_golden_render.py -> tests/golden/render_rays.npz)."""
import numpy as np
import pytest
import torch

from conftest import golden_decoder


def test_oracle_render_matches_reference_golden(golden):
    from oracle import field_oracle as fo

    g = golden("render_rays.npz")
    ws, bs = golden_decoder(g)
    t_near, t_far, valid = fo.rays_intersect_bbox(g["rays_o"], g["rays_d"], float(g["radius"]))
    np.testing.assert_array_equal(valid, g["rays_valid"])
    np.testing.assert_array_equal(t_near, g["t_near"][:, 0])
    np.testing.assert_array_equal(t_far, g["t_far"][:, 0])
    comp = fo.render_rays(g["triplane"], g["rays_o"].reshape(15, 20, 3), g["rays_d"].reshape(15, 20, 3), ws, bs,
                          radius=float(g["radius"]), num_samples=int(g["num_samples"]))
    assert comp.shape == (15, 20, 3)
    assert np.abs(comp - g["comp_rgb"]).max() < 2e-6
    assert g["comp_rgb"].min() < 0.6 and g["comp_rgb"].max() > 0.95  # opaque and nearly transparent rays both present


def test_host_rays_intersect_bbox_matches_reference_golden(golden):
    from sculptmate_b200.tsr.utils import rays_intersect_bbox

    g = golden("render_rays.npz")
    o, d = torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"])
    t_near, t_far, valid = rays_intersect_bbox(o.view(15, 20, 3), d.view(15, 20, 3), float(g["radius"]))
    assert t_near.shape == (15, 20, 1) and valid.shape == (15, 20)
    np.testing.assert_array_equal(t_near.reshape(-1, 1).numpy(), g["t_near"])
    np.testing.assert_array_equal(t_far.reshape(-1, 1).numpy(), g["t_far"])
    # a ray pointing away from the box is invalid and gets t = 0 like the reference (utils.py:140-143)
    o2 = torch.tensor([[2.0, 0.0, 0.0]])
    tn, tf, v = rays_intersect_bbox(o2, o2 / 2, 0.87)
    assert not bool(v[0]) and float(tn) == 0.0 and float(tf) == 0.0


def _renderer_and_decoder(g):
    from test_gpu_field import _model

    m = _model(g)
    return m.renderer, m.decoder


@pytest.mark.gpu
def test_sample_positions_bit_exact_and_composite(golden):
    import ctypes

    from sculptmate_b200 import _capi
    from sculptmate_b200.tsr.utils import rays_intersect_bbox

    g = golden("render_rays.npz")
    lib = _capi.load()
    o, d = torch.from_numpy(g["rays_o"]).cuda(), torch.from_numpy(g["rays_d"]).cuda()
    S = 128
    t_near, t_far, _ = rays_intersect_bbox(o, d, 0.87)
    t_vals = torch.linspace(0, 1, S + 1)
    t_mid = ((t_vals[:-1] + t_vals[1:]) / 2.0).cuda()
    xyz = torch.empty((300, S, 3), device="cuda")
    assert lib.smb_ray_sample_positions(o.data_ptr(), d.data_ptr(), t_near.data_ptr(), t_far.data_ptr(), t_mid.data_ptr(), 300, S, xyz.data_ptr(), None) == 0
    z = t_near * (1 - t_mid[None]) + t_far * t_mid[None]
    ref = o[:, None, :] + z[..., None] * d[..., None, :]  # nerf_renderer.py:113-117
    torch.cuda.synchronize()
    assert torch.equal(xyz, ref)
    # composite on random inputs, ragged sample counts (not a multiple of 32) included
    for S2 in (1, 31, 128, 200):
        gen = torch.Generator().manual_seed(S2)
        sigma = (torch.rand(77, S2, generator=gen) * 40).cuda()
        col = torch.rand(77, S2, 3, generator=gen).cuda()
        deltas = torch.full((S2,), 1.0 / S2).cuda()
        comp = torch.empty(77, 3, device="cuda")
        op = torch.empty(77, device="cuda")
        assert lib.smb_ray_composite(sigma.data_ptr(), col.data_ptr(), deltas.data_ptr(), 77, S2, comp.data_ptr(), op.data_ptr(), None) == 0
        alpha = 1 - torch.exp(-deltas * sigma)
        acc = torch.cat([torch.ones_like(alpha[:, :1]), torch.cumprod(1 - alpha[:, :-1] + 1e-10, dim=-1)], dim=-1)
        w = alpha * acc
        ref_rgb = (w[..., None] * col).sum(dim=-2) + (1 - w.sum(dim=-1))[:, None]
        torch.cuda.synchronize()
        assert (comp - ref_rgb).abs().max() < 5e-6 and (op - w.sum(dim=-1)).abs().max() < 5e-6


@pytest.mark.gpu
def test_forward_dropin_vs_reference_golden(golden):
    g = golden("render_rays.npz")
    rend, dec = _renderer_and_decoder(g)
    tp = torch.from_numpy(g["triplane"]).cuda()
    o = torch.from_numpy(g["rays_o"]).cuda().view(15, 20, 3)
    d = torch.from_numpy(g["rays_d"]).cuda().view(15, 20, 3)
    ref = g["comp_rgb"]
    c32 = rend(dec, tp, o, d, precision="fp32")
    assert c32.shape == (15, 20, 3) and c32.dtype == torch.float32
    assert np.abs(c32.cpu().numpy() - ref).max() < 1e-5  # fp32 field query: reference precision
    ctc = rend(dec, tp, o, d)  # default: tensor-core query (fp16 operands, fp32 accumulate)
    # the golden decoder's density row is scaled x40 (oracle/make_golden_render.py), so the fp16-operand logit error
    # (~1e-3 abs at unit scale) becomes ~4e-2 in log-density; on the composite that stays below 2e-2
    assert np.abs(ctc.cpu().numpy() - ref).max() < 2e-2
    # batched scene codes: one _forward per scene (nerf_renderer.py:164-171)
    tp2 = torch.stack([tp, tp.flip(-1)])
    ob = torch.stack([o.view(-1, 3)[:50], o.view(-1, 3)[50:100]])
    db = torch.stack([d.view(-1, 3)[:50], d.view(-1, 3)[50:100]])
    cb = rend(dec, tp2, ob, db, precision="fp32")
    assert np.abs(cb.cpu().numpy() - g["comp_rgb_batched"]).max() < 1e-5


@pytest.mark.gpu
def test_forward_raises_like_reference_when_a_ray_misses(golden):
    g = golden("render_rays.npz")
    rend, dec = _renderer_and_decoder(g)
    tp = torch.from_numpy(g["triplane"]).cuda()
    o = torch.from_numpy(g["rays_o"]).cuda()
    d = torch.from_numpy(g["rays_d"]).cuda().clone()
    d[7] = o[7] / o[7].norm()  # points away from the box
    with pytest.raises(RuntimeError):
        rend(dec, tp, o, d)
