"""Live re-validation of the oracle and of the CPU port against the UNMODIFIED reference,
where /root/reference exists (skipped on the GPU box, which only has the committed goldens)."""
import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference is not present on this machine")
RADIUS = 0.87


def test_query_triplane_restatement_vs_live_reference():
    from oracle import field_oracle as fo

    ref = ref_shim.load_triposr()
    dec = ref_shim.make_reference_decoder(5)
    rend = ref_shim.make_reference_renderer(1000)
    torch.manual_seed(6)
    tp = torch.randn(3, 40, 20, 20)
    pos = (torch.rand(3000, 3) * 2 - 1) * 1.05 * RADIUS  # a few % outside the radius: zero padding
    with torch.no_grad():
        out = rend.query_triplane(dec, pos, tp)
    ws, bs = fo.decoder_params_from_state_dict({k: v.numpy() for k, v in dec.state_dict().items()})
    mine = fo.query_triplane(pos.numpy(), tp.numpy(), ws, bs, radius=RADIUS)
    for k in ("density", "features", "density_act", "color"):
        assert np.abs(mine[k] - out[k].numpy()).max() < 2e-5, k


def test_cpu_port_matches_live_reference_extract_mesh_inputs():
    """oracle/cpu_reference_port.py issues the same aten calls as the reference: same density."""
    from oracle import cpu_reference_port as port

    ref = ref_shim.load_triposr()
    dec = ref_shim.make_reference_decoder(2)
    rend = ref_shim.make_reference_renderer(8192)
    torch.manual_seed(3)
    tp = torch.randn(3, 40, 16, 16)
    R = 20
    h = ref.isosurface.MarchingCubeHelper(R)
    pos = ref.utils.scale_tensor(h.grid_vertices, (0, 1), (-RADIUS, RADIUS))
    with torch.no_grad():
        want = rend.query_triplane(dec, pos, tp)["density_act"]
    sd = dec.state_dict()
    layers = port.make_layers([sd[f"layers.{i}.weight"].numpy() for i in range(0, 20, 2)], [sd[f"layers.{i}.bias"].numpy() for i in range(0, 20, 2)])
    got = port.query_triplane(layers, pos, tp)["density_act"]
    assert torch.equal(got, want)  # identical calls, identical bits


def test_marching_tets_restatement_vs_live_reference(tmp_path):
    from oracle import sf3d_oracle as so
    from sculptmate_b200.sf3d.tets import kuhn_tet_grid, save_tet_grid

    ref = ref_shim.load_sf3d()
    n = 7
    helper = ref.isosurface.MarchingTetrahedraHelper(n, save_tet_grid(str(tmp_path / "t.npz"), n))
    g = torch.Generator().manual_seed(1)
    sdf = torch.randn(helper.grid_vertices.shape[0], 1, generator=g)
    with torch.no_grad():
        mesh = helper(sdf, None)
    v, f = so.marching_tets(*kuhn_tet_grid(n)[:1], sdf.numpy(), kuhn_tet_grid(n)[1])
    np.testing.assert_array_equal(f, mesh.t_pos_idx.numpy())
    np.testing.assert_array_equal(v, mesh.v_pos.numpy())


@pytest.mark.parametrize("seed,res", [(11, 24), (12, 33), (13, 40)])
def test_bake_restatement_vs_live_reference_python_functions(seed, res):
    """oracle/bake_oracle.py against the reference's rasterize_cpu / interpolate_cpu (texture_baker/common.py) on fresh
    random atlases: same mask and -- wherever the same triangle is chosen -- bit-identical barycentrics and attributes."""
    from oracle import bake_oracle as bo
    from oracle.make_golden_bake import atlas, load_common

    tb = load_common()
    uv, faces, attr = atlas(seed, n=6)
    with np.errstate(all="ignore"):
        ref_rast = tb.rasterize_cpu(uv, faces, res).astype(np.float32)
        ref_int = tb.interpolate_cpu(attr, faces, ref_rast).astype(np.float32)
    rast = bo.rasterize(uv, faces, res)
    np.testing.assert_array_equal(rast[..., 3] >= 0, ref_rast[..., 3] >= 0)
    same = rast[..., 3] == ref_rast[..., 3]
    assert same.mean() > 0.995  # only texels exactly on a shared edge may pick the other triangle (BVH order vs lowest index)
    np.testing.assert_array_equal(rast[same], ref_rast[same])
    out = bo.interpolate(attr, faces, rast)
    np.testing.assert_array_equal(out[same], ref_int[same])
    assert np.abs(out - ref_int).max() < 1e-5


def test_render_restatement_vs_live_reference():
    """oracle/field_oracle.py::render_rays against the reference's TriplaneNeRFRenderer.forward on fresh rays."""
    from oracle import field_oracle as fo
    from oracle.make_golden_render import make_rays

    ref_shim.load_triposr()
    dec = ref_shim.make_reference_decoder(9)
    with torch.no_grad():
        dec.layers[18].weight[0] *= 30.0
        dec.layers[18].bias[0] -= 2.0
    rend = ref_shim.make_reference_renderer(4096)
    rend.eval()
    torch.manual_seed(4)
    tp = 0.5 * torch.randn(3, 40, 12, 12)
    rays_o, rays_d = make_rays(64, 33)
    with torch.no_grad():
        want = rend(dec, tp, rays_o, rays_d).numpy()
    ws, bs = fo.decoder_params_from_state_dict({k: v.numpy() for k, v in dec.state_dict().items()})
    got = fo.render_rays(tp.numpy(), rays_o.numpy(), rays_d.numpy(), ws, bs, radius=RADIUS, num_samples=int(rend.cfg.num_samples_per_ray))
    assert np.abs(got - want).max() < 5e-6
    assert want.min() < 0.9  # not all background


def test_rays_intersect_bbox_mirror_vs_live_reference():
    """sculptmate_b200.tsr.utils.rays_intersect_bbox against the reference's (tsr/utils.py:115-149) on random rays that hit,
    graze and miss the box, incl. zero direction components: bit-identical t_near / t_far / validity."""
    from sculptmate_b200.tsr.utils import rays_intersect_bbox

    ref = ref_shim.load_triposr()
    g = torch.Generator().manual_seed(8)
    o = torch.randn(4000, 3, generator=g) * 1.5
    d = torch.randn(4000, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    d[:50, 0] = 0.0
    d[50:100, 1:] = 0.0
    want = ref.utils.rays_intersect_bbox(o.clone(), d.clone(), RADIUS)
    got = rays_intersect_bbox(o.clone(), d.clone(), RADIUS)
    assert 0.1 < float(want[2].float().mean()) < 0.9
    for a, b in zip(got, want):
        assert a.shape == b.shape and torch.equal(a, b)
    # batched shapes
    got2 = rays_intersect_bbox(o.view(40, 100, 3), d.view(40, 100, 3), RADIUS)
    want2 = ref.utils.rays_intersect_bbox(o.view(40, 100, 3), d.view(40, 100, 3), RADIUS)
    for a, b in zip(got2, want2):
        assert a.shape == b.shape and torch.equal(a, b)


def test_sf3d_activation_table_vs_live_reference():
    """Every name of the reference's get_activation (sf3d/models/network.py:98-136) gives the same values through the
    drop-in's table, including normalize's dtype-dependent eps (models/utils.py:57-76) -- ADVICE round 1."""
    from sculptmate_b200.sf3d.models import network as mine

    ref = ref_shim.load_sf3d()
    import importlib

    ref_net = importlib.import_module("sf3d.models.network")
    torch.manual_seed(0)
    x = torch.randn(64, 7, 3) * 2
    x[0] = 0.0  # zero vectors: the eps of normalize decides the result
    x[1] = 1e-9
    names = ["none", "linear", "identity", "lin2srgb", "exp", "shifted_exp", "trunc_exp", "shifted_trunc_exp", "sigmoid", "tanh",
             "shifted_softplus", "scale_-11_01", "negative", "normalize_channel_last", "normalize_channel_first", "relu", "silu", None]
    for n in names:
        a, b = mine.get_activation(n)(x), ref_net.get_activation(n)(x)
        assert torch.equal(a, b), n
    assert torch.equal(mine.get_activation("normalize_channel_last")(x.half()), ref_net.get_activation("normalize_channel_last")(x.half()))
    with pytest.raises(ValueError):
        mine.get_activation("no_such_activation")
    del ref
