"""GPU parity of the field kernels (query_triplane + NeRFMLP) through the C ABI.

Tolerances (stated, per north_star): the fp32 CUDA-core kernel must agree with the
reference within fp32 reimplementation noise; the tensor-core lattice kernel feeds fp16
operands to tcgen05.mma with fp32 accumulation and uses tanh.approx for SiLU, so it is
held to  max|dlogit| <= 1e-2,  mean|dlogit| <= 2e-4,  max rel err of density_act <= 1e-2
on N(0,1) planes (logit span ~6) -- about 2.5x what is measured -- and to 3e-3 on the baked field the benchmark uses.
"""
import hashlib

import numpy as np
import pytest
import torch

from conftest import golden_decoder

pytestmark = pytest.mark.gpu

RADIUS = 0.87
# measured (round 2): max |dlogit| 4.2e-3 on N(0,1) planes (logit span ~6; worst over R = 8..256), mean 5.6e-5; 8.9e-4 on the
# baked bench field.  The bounds are 2.5-3x the measurement so that a numerical regression (e.g. a worse activation
# approximation) fails the suite: 1e-2 on N(0,1) planes, 3e-3 on the baked field (round 1 asserted 2e-2 for both).
TC_MAX_ABS, TC_MEAN_ABS, TC_MAX_REL = 1e-2, 2e-4, 1e-2
TC_BAKED_MAX_REL = 3e-3


def _pack(g):
    from sculptmate_b200 import runtime

    ws, bs = golden_decoder(g)
    blob, lay = runtime.pack_decoder_host([torch.from_numpy(w) for w in ws], [torch.from_numpy(b) for b in bs])
    return ws, bs, runtime.DecoderPack(blob.cuda(), lay, ("test",))


def _model(g):
    from sculptmate_b200.tsr import TSR

    ws, bs = golden_decoder(g)
    m = TSR()
    sd = {}
    for i in range(10):
        sd[f"layers.{2 * i}.weight"] = torch.from_numpy(ws[i])
        sd[f"layers.{2 * i}.bias"] = torch.from_numpy(bs[i])
    m.decoder.load_state_dict(sd)
    return m.cuda()


def _triplane64(g):
    torch.manual_seed(int(g["triplane_seed"]))
    tp = torch.randn(3, 40, 64, 64)
    assert hashlib.sha256(tp.numpy().tobytes()).digest() == bytes(g["triplane_sha256"].tobytes())
    return tp


@pytest.mark.parametrize("name", ["field_small.npz", "field_64.npz"])
def test_query_points_fp32_matches_reference_golden(golden, name):
    from sculptmate_b200 import runtime

    g = golden(name)
    _, _, pack = _pack(g)
    tp = torch.from_numpy(g["triplane"]) if "triplane" in g else _triplane64(g)
    scene = runtime.prepare_scene(tp.cuda(), pack)
    out = runtime.query_points(scene, pack, torch.from_numpy(g["positions"]).cuda(), RADIUS, -1.0)
    for k in ("density", "features", "density_act", "color"):
        assert out[k].shape == g[k].shape and out[k].dtype == torch.float32
        assert np.abs(out[k].cpu().numpy() - g[k]).max() < 2e-5, k  # fp32 noise floor (SURVEY: 1.2e-5)


def test_query_triplane_dropin_shapes_and_values(golden):
    g = golden("field_small.npz")
    m = _model(g)
    tp = torch.from_numpy(g["triplane"]).cuda()
    pos = torch.from_numpy(g["positions"][:240]).cuda().view(4, 6, 10, 3)
    out = m.renderer.query_triplane(m.decoder, pos, tp)
    assert set(out) == {"density", "features", "density_act", "color"}
    assert out["density"].shape == (4, 6, 10, 1) and out["color"].shape == (4, 6, 10, 3)
    assert np.abs(out["density_act"].cpu().numpy().reshape(-1, 1) - g["density_act"][:240]).max() < 2e-5
    e = m.renderer.query_triplane(m.decoder, torch.zeros(0, 3).cuda(), tp)  # empty input
    assert e["density"].shape == (0, 1)
    m.renderer.set_chunk_size(8192)  # harmless knob
    out2 = m.renderer.query_triplane(m.decoder, pos, tp)
    assert torch.equal(out2["density"], out["density"])


def test_decoder_pack_refreshes_on_weight_change(golden):
    g = golden("field_small.npz")
    m = _model(g)
    tp = torch.from_numpy(g["triplane"]).cuda()
    pos = torch.from_numpy(g["positions"][:64]).cuda()
    a = m.renderer.query_triplane(m.decoder, pos, tp)["density"].clone()
    with torch.no_grad():
        m.decoder.layers[18].bias.add_(1.0)
    b = m.renderer.query_triplane(m.decoder, pos, tp)["density"]
    assert torch.allclose(b, a + 1.0, atol=1e-5)


@pytest.mark.parametrize("R", [8, 32, 64, 100, 128, 160])
def test_lattice_tc_vs_fp32(golden, R):
    """Same lattice, same coordinates: tensor-core kernel vs fp32 kernel."""
    from sculptmate_b200 import runtime

    g = golden("field_64.npz")
    _, _, pack = _pack(g)
    scene = runtime.prepare_scene(_triplane64(g).cuda(), pack)
    ax = runtime.lattice_axis(R, RADIUS, device="cuda")
    act32, raw32 = runtime.query_lattice(scene, pack, ax, R, RADIUS, -1.0, precision="fp32", want_raw=True)
    act, raw = runtime.query_lattice(scene, pack, ax, R, RADIUS, -1.0, precision="tc", want_raw=True)
    assert act.shape == (R, R, R) and not torch.isnan(act).any()
    d = (raw - raw32).abs()
    rel = (act / act32 - 1).abs().max().item()
    print(f"R={R} max|dlogit|={d.max().item():.2e} mean={d.mean().item():.2e} max_rel(density_act)={rel:.2e}")
    assert d.max().item() <= TC_MAX_ABS and d.mean().item() <= TC_MEAN_ABS and rel <= TC_MAX_REL
    # slab evaluation reproduces the same bits (sharding must not change the result)
    a, n = R // 3, max(1, R // 2)
    sl = runtime.query_lattice(scene, pack, ax, R, RADIUS, -1.0, x_begin=a, nx=n, precision="tc")
    assert torch.equal(sl, act[a : a + n])
    again = runtime.query_lattice(scene, pack, ax, R, RADIUS, -1.0, precision="tc")
    assert torch.equal(again, act)  # deterministic


def test_lattice_vs_cpu_oracle(golden):
    from oracle import field_oracle as fo
    from sculptmate_b200 import runtime

    g = golden("field_64.npz")
    ws, bs, pack = _pack(g)
    tp = _triplane64(g)
    scene = runtime.prepare_scene(tp.cuda(), pack)
    R = 24
    ax = runtime.lattice_axis(R, RADIUS, device="cuda")
    ref = fo.grid_density(R, tp.numpy(), ws, bs)
    a32 = runtime.query_lattice(scene, pack, ax, R, RADIUS, -1.0, precision="fp32").cpu().numpy()
    atc = runtime.query_lattice(scene, pack, ax, R, RADIUS, -1.0, precision="tc").cpu().numpy()
    # density_act = exp(logit): its relative error is the logit's absolute error; two fp32
    # evaluations of a 10-layer MLP in different summation orders differ by a few 1e-5
    assert np.abs(a32 / ref - 1).max() < 1e-4
    assert np.abs(atc / ref - 1).max() < TC_MAX_REL


def test_lattice_reference_golden_density(golden):
    """The reference's own density grid from its extract_mesh body (golden)."""
    g = golden("extract_mesh.npz")
    m = _model(g)
    R = int(g["resolution"])
    tp = torch.from_numpy(g["triplane"]).cuda()
    d32 = m.renderer.query_lattice(m.decoder, tp, R, precision="fp32").cpu().numpy()
    dtc = m.renderer.query_lattice(m.decoder, tp, R, precision="tc").cpu().numpy()
    assert np.abs(d32 / g["density_act"] - 1).max() < 2e-5
    # baked field (what the benchmark uses): small logit span -> 5x tighter
    assert np.abs(dtc / g["density_act"] - 1).max() < TC_BAKED_MAX_REL


def test_full_size_256_sampled_planes(golden):
    """BASELINE size (256^3): tensor-core result checked on sampled x-planes against the
    fp32 kernel, plus determinism of the full grid."""
    from sculptmate_b200 import runtime

    g = golden("field_64.npz")
    _, _, pack = _pack(g)
    scene = runtime.prepare_scene(_triplane64(g).cuda(), pack)
    R = 256
    ax = runtime.lattice_axis(R, RADIUS, device="cuda")
    full, raw = runtime.query_lattice(scene, pack, ax, R, RADIUS, -1.0, want_raw=True)
    assert torch.isfinite(full).all()
    for x0 in (0, 97, 254):
        a32, r32 = runtime.query_lattice(scene, pack, ax, R, RADIUS, -1.0, x_begin=x0, nx=2, precision="fp32", want_raw=True)
        d = (raw[x0 : x0 + 2] - r32).abs()
        assert d.max().item() <= TC_MAX_ABS and d.mean().item() <= TC_MEAN_ABS
        assert (full[x0 : x0 + 2] / a32 - 1).abs().max().item() <= TC_MAX_REL
    assert torch.equal(runtime.query_lattice(scene, pack, ax, R, RADIUS, -1.0), full)


@pytest.mark.parametrize("name", ["field_small.npz", "field_64.npz"])
def test_query_points_tensor_core_vs_reference_golden(golden, name):
    """tcgen05 points kernel (fp16 operands, fp32 accumulate, tanh.approx) vs the reference's fp32
    outputs at arbitrary positions, incl. border / out-of-range positions (zero padding)."""
    g = golden(name)
    m = _model(g)
    tp = torch.from_numpy(g["triplane"]).cuda() if "triplane" in g.files else _triplane64(g).cuda()
    pos = torch.from_numpy(g["positions"]).cuda()
    out = m.renderer.query_triplane(m.decoder, pos, tp, precision="tc")
    ref32 = m.renderer.query_triplane(m.decoder, pos, tp, precision="fp32")
    for k in ("density", "features", "density_act", "color"):
        assert out[k].shape == g[k].shape and out[k].dtype == torch.float32
    assert np.abs(out["density"].cpu().numpy() - g["density"]).max() < TC_MAX_ABS
    assert np.abs(out["density"].cpu().numpy() - g["density"]).mean() < 10 * TC_MEAN_ABS
    assert np.abs(out["features"].cpu().numpy() - g["features"]).max() < TC_MAX_ABS
    assert np.abs(out["density_act"].cpu().numpy() / g["density_act"] - 1).max() < TC_MAX_REL
    assert np.abs(out["color"].cpu().numpy() - g["color"]).max() < 5e-3
    assert torch.equal(m.renderer.query_triplane(m.decoder, pos, tp, precision="tc")["density"], out["density"])  # deterministic
    assert (out["density"] - ref32["density"]).abs().max() < TC_MAX_ABS
    ragged = m.renderer.query_triplane(m.decoder, pos[:131], tp, precision="tc")  # one full tile + 3 rows
    assert torch.equal(ragged["density"], out["density"][:131])


def test_nerfmlp_forward_on_features(golden):
    """NeRFMLP.forward (network_utils.py:116-124) on pre-computed features runs in the CUDA library."""
    from oracle import field_oracle as fo

    g = golden("field_small.npz")
    m = _model(g)
    ws, bs = golden_decoder(g)
    rng = np.random.RandomState(0)
    x = rng.randn(3, 70, 120).astype(np.float32)
    out = m.decoder(torch.from_numpy(x).cuda())
    ref = fo.nerf_mlp(x.reshape(-1, 120), ws, bs)
    assert out["density"].shape == (3, 70, 1) and out["features"].shape == (3, 70, 3)
    assert np.abs(out["density"].cpu().numpy().reshape(-1, 1) - ref["density"]).max() < 2e-5
    assert np.abs(out["features"].cpu().numpy().reshape(-1, 3) - ref["features"]).max() < 2e-5


def test_full_size_256_planes_vs_cpu_oracle(golden):
    """256^3 tensor-core lattice against the CPU ORACLE itself (numpy fp32 restatement of query_triplane + NeRFMLP, pinned
    to the reference's goldens), not against another CUDA kernel: three whole x-planes incl. both lattice borders."""
    from oracle import field_oracle as fo
    from sculptmate_b200 import runtime

    g = golden("field_64.npz")
    ws, bs, pack = _pack(g)
    tp = _triplane64(g)
    scene = runtime.prepare_scene(tp.cuda(), pack)
    R = 256
    ax = runtime.lattice_axis(R, RADIUS, device="cuda")
    full, raw = runtime.query_lattice(scene, pack, ax, R, RADIUS, -1.0, want_raw=True)
    axis = fo.grid_axis(R)
    for i in (0, 131, 255):
        x, y, z = np.meshgrid(axis[i : i + 1], axis, axis, indexing="ij")
        pos = fo.scale_tensor(np.stack([x.reshape(-1), y.reshape(-1), z.reshape(-1)], -1), (0, 1), (-RADIUS, RADIUS))
        ref = fo.query_triplane(pos, tp.numpy(), ws, bs, radius=RADIUS)
        d = np.abs(raw[i].cpu().numpy().reshape(-1) - (ref["density"].reshape(-1)))
        rel = np.abs(full[i].cpu().numpy().reshape(-1) / ref["density_act"].reshape(-1) - 1)
        print(f"plane {i}: max|dlogit| {d.max():.2e} mean {d.mean():.2e} max rel {rel.max():.2e}")
        assert d.max() <= TC_MAX_ABS and d.mean() <= TC_MEAN_ABS and rel.max() <= TC_MAX_REL
