"""GPU parity of the Stable Fast 3D mesh path (csrc/sf3d.cu through the C ABI) against the
fixtures produced by the unmodified reference and against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import sf3d_oracle as so
from sculptmate_b200.sf3d.tets import kuhn_tet_grid, save_tet_grid

pytestmark = pytest.mark.gpu
RADIUS = 0.87


def _sd(g):
    return {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}


@pytest.fixture(scope="module")
def tets_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("tets")
    for n in (10, 12, 40):
        save_tet_grid(os.path.join(d, f"tets{n}.npz"), n)
    return str(d)


def _model(g, tets_dir, n, thr, precision="fp32"):
    from sculptmate_b200.sf3d import SF3D

    m = SF3D(dict(isosurface_resolution=n, isosurface_threshold=thr, radius=RADIUS, precision=precision,
                  tets_path=os.path.join(tets_dir, f"tets{n}.npz")))
    m.decoder.load_state_dict(_sd(g))
    return m.cuda()


@pytest.mark.parametrize("name", ["sphere", "torus", "noise"])
def test_marching_tets_bit_exact_vs_reference(golden, tets_dir, name):
    from sculptmate_b200.sf3d import MarchingTetrahedraHelper

    g = golden("sf3d_mtet.npz")
    n = int(g["n"])
    h = MarchingTetrahedraHelper(n, os.path.join(tets_dir, f"tets{n}.npz")).cuda()
    np.testing.assert_array_equal(h.all_edges.cpu().numpy(), g["all_edges"])
    level = torch.from_numpy(g[f"{name}_sdf"]).cuda().view(-1, 1)
    deform = torch.from_numpy(g[f"{name}_deform"]).cuda() if f"{name}_deform" in g.files else None
    mesh = h(level, deform)
    assert mesh.v_pos.dtype == torch.float32 and mesh.t_pos_idx.dtype == torch.int64 and mesh.v_pos.is_cuda
    np.testing.assert_array_equal(mesh.t_pos_idx.cpu().numpy(), g[f"{name}_f"])
    if deform is None:
        np.testing.assert_array_equal(mesh.v_pos.cpu().numpy(), g[f"{name}_v"])  # same fp32 operation order
    else:  # tanhf (CUDA) vs the reference's CPU tanh: <= 1 ulp on the deformed grid
        assert np.abs(mesh.extras["grid_vertices"].cpu().numpy() - g[f"{name}_grid"]).max() < 2e-7
        assert np.abs(mesh.v_pos.cpu().numpy() - g[f"{name}_v"]).max() < 1e-5
    assert set(mesh.extras) == {"grid_vertices", "tet_edges", "grid_level", "grid_deformation"}


def test_query_triplane_and_heads_vs_reference(golden, tets_dir):
    g = golden("sf3d_path.npz")
    m = _model(g, tets_dir, int(g["n"]), float(g["threshold"]))
    pos = torch.from_numpy(g["positions"]).cuda()
    tp = torch.from_numpy(g["triplane"]).cuda()
    feats = m.query_triplane(pos, tp)
    assert feats.shape == (1, pos.shape[0], 120)  # the reference keeps the batch dim it adds
    assert np.abs(feats.cpu().numpy() - g["features"]).max() < 2e-6
    fb = m.query_triplane(pos[None].repeat(2, 1, 1), tp[None].repeat(2, 1, 1, 1, 1))
    assert fb.shape == (2, pos.shape[0], 120) and torch.equal(fb[0], feats[0]) and torch.equal(fb[1], feats[0])
    dec = m.decoder(feats, include=["vertex_offset", "density"])
    assert dec["density"].shape == (1, pos.shape[0], 1) and dec["vertex_offset"].shape == (1, pos.shape[0], 3)
    assert np.abs(dec["density"].cpu().numpy() / g["density"] - 1).max() < 2e-5
    assert np.abs(dec["vertex_offset"].cpu().numpy() - g["vertex_offset"]).max() < 2e-5
    only = m.decoder(feats, exclude=["vertex_offset"])
    assert list(only) == ["density"] and torch.equal(only["density"], dec["density"])


def test_triplane_to_meshes_vs_reference(golden, tets_dir):
    g = golden("sf3d_path.npz")
    n = int(g["n"])
    m = _model(g, tets_dir, n, float(g["threshold"]))
    tp = torch.from_numpy(g["triplane"]).cuda()
    meshes = m.triplane_to_meshes(tp[None].repeat(2, 1, 1, 1, 1))
    assert len(meshes) == 2
    mesh = meshes[0]
    level = mesh.extras["grid_level"].cpu().numpy().reshape(-1)
    flips = (np.sign(level) != np.sign(g["grid_level"].reshape(-1))).mean()
    assert flips < 1e-3
    assert np.abs(mesh.extras["grid_vertices"].cpu().numpy() - g["grid_vertices"]).max() < 1e-6
    if flips == 0:
        np.testing.assert_array_equal(mesh.t_pos_idx.cpu().numpy(), g["t_pos_idx"])
        assert np.abs(mesh.v_pos.cpu().numpy() - g["v_pos"]).max() < 5e-4
    assert torch.equal(meshes[1].t_pos_idx, mesh.t_pos_idx) and torch.equal(meshes[1].v_pos, mesh.v_pos)
    # and bit-exact against the oracle fed the GPU's own level / deformed grid
    _, tets = kuhn_tet_grid(n)
    v, f = so.marching_tets(mesh.extras["grid_vertices"].cpu().numpy(), level, tets)
    np.testing.assert_array_equal(mesh.t_pos_idx.cpu().numpy(), f)
    v = (v * np.float32(2 * RADIUS) + np.float32(-RADIUS)).astype(np.float32)
    np.testing.assert_array_equal(mesh.v_pos.cpu().numpy(), v)


def test_larger_grid_properties(tets_dir):
    """n = 40 (68 921 vertices, 384 000 tets): closed surface, Euler characteristic 2."""
    from conftest import mesh_topology
    from sculptmate_b200.sf3d import MarchingTetrahedraHelper

    n = 40
    h = MarchingTetrahedraHelper(n, os.path.join(tets_dir, f"tets{n}.npz")).cuda()
    gv = h.grid_vertices
    sdf = (0.33 - (gv - 0.5).norm(dim=-1)).view(-1, 1)
    mesh = h(sdf, None)
    v, f = mesh.v_pos.cpu().numpy(), mesh.t_pos_idx.cpu().numpy()
    closed, chi, vol = mesh_topology(v, f)
    assert closed and chi == 2 and abs(abs(vol) - 4 / 3 * np.pi * 0.33**3) < 2e-3
    again = h(sdf, None)
    assert torch.equal(again.v_pos, mesh.v_pos) and torch.equal(again.t_pos_idx, mesh.t_pos_idx)  # deterministic
    # empty surface: no vertices, no faces, no exception (the reference returns empty tensors)
    none = h(torch.ones_like(sdf), None)
    assert none.v_pos.shape == (0, 3) and none.t_pos_idx.shape == (0, 3)


def test_tensor_core_query_vs_fp32_and_mesh(golden, tets_dir):
    """precision="tc" (tcgen05, fp16 operands): density / offsets within the stated tolerance of the
    fp32 path, and the mesh it yields is the oracle's mesh for the level it produced."""
    from sculptmate_b200 import runtime

    g = golden("sf3d_path.npz")
    n = int(g["n"])
    m = _model(g, tets_dir, n, float(g["threshold"]), precision="tc")
    tp = torch.from_numpy(g["triplane"]).cuda()
    pos = torch.from_numpy(g["positions"]).cuda()
    planes = runtime.prepare_planes_cl(tp)
    r = runtime.query_points_tc(planes, runtime.get_sf3d_points_pack(m.decoder, pos.device), pos, RADIUS, -1.0,
                                align_corners=True, sigmoid_vec=False, want=("out0_act", "vec"))
    assert np.abs(r["out0_act"].cpu().numpy() / g["density"][0] - 1).max() < 2e-2  # stated fp16-operand tolerance
    assert np.abs(r["vec"].cpu().numpy() - g["vertex_offset"][0]).max() < 2e-2
    rh = runtime.query_points_tc(runtime.prepare_planes_half(tp), runtime.get_sf3d_points_pack(m.decoder, pos.device), pos, RADIUS, -1.0,
                                 align_corners=True, sigmoid_vec=False, want=("out0_act", "vec"))  # fp16 planes (the default of "tc")
    assert np.abs(rh["out0_act"].cpu().numpy() / g["density"][0] - 1).max() < 2e-2
    assert np.abs(rh["vec"].cpu().numpy() - g["vertex_offset"][0]).max() < 2e-2
    mesh = m.triplane_to_meshes(tp[None])[0]
    level = mesh.extras["grid_level"].cpu().numpy().reshape(-1)
    assert (np.sign(level) != np.sign(g["grid_level"].reshape(-1))).mean() < 2e-2
    _, tets = kuhn_tet_grid(n)
    v, f = so.marching_tets(mesh.extras["grid_vertices"].cpu().numpy(), level, tets)
    np.testing.assert_array_equal(mesh.t_pos_idx.cpu().numpy(), f)
    v = (v * np.float32(2 * RADIUS) + np.float32(-RADIUS)).astype(np.float32)
    np.testing.assert_array_equal(mesh.v_pos.cpu().numpy(), v)


def test_texel_query_heads_on_tensor_cores(tets_dir):
    """Texel-space query of the texture bake (sf3d/system.py:375-378): query_triplane at the baked positions +
    decoder(exclude=[density, vertex_offset]) -> heads ``features`` (sigmoid) and ``perturb_normal`` (normalised),
    fused per head on the tensor cores, against the CPU oracle (oracle/field_oracle.py, pinned to the reference)."""
    from oracle import field_oracle as fo
    from sculptmate_b200.sf3d import SF3D

    heads = [  # StableFast/checkpoints/config.yaml:49-65
        dict(name="density", out_channels=1, out_bias=-1.0, n_hidden_layers=2, output_activation="trunc_exp"),
        dict(name="features", out_channels=3, n_hidden_layers=3, output_activation="sigmoid"),
        dict(name="perturb_normal", out_channels=3, n_hidden_layers=3, output_activation="normalize_channel_last"),
        dict(name="vertex_offset", out_channels=3, n_hidden_layers=2),
    ]
    torch.manual_seed(5)
    m = SF3D(dict(isosurface_resolution=10, radius=RADIUS, tets_path=os.path.join(tets_dir, "tets10.npz"),
                  decoder=dict(in_channels=120, n_neurons=64, activation="silu", heads=heads))).cuda()
    g = torch.Generator().manual_seed(6)
    tp = (0.5 * torch.randn(3, 40, 48, 48, generator=g)).cuda()
    pos = ((torch.rand(3001, 3, generator=g) * 2 - 1) * RADIUS).cuda()
    sd = {k: v.detach().cpu().numpy() for k, v in m.decoder.state_dict().items()}
    feats = fo.sf3d_query_triplane(pos.cpu().numpy(), tp.cpu().numpy(), RADIUS)

    def ref(name, n_lin, act):
        ws = [sd[f"heads.{name}.{2 * i}.weight"] for i in range(n_lin)]
        bs = [sd[f"heads.{name}.{2 * i}.bias"] for i in range(n_lin)]
        return fo.material_mlp_head(feats, ws, bs, 0.0, act)

    r_feat = ref("features", 4, "sigmoid")
    r_nrm = ref("perturb_normal", 4, None)
    r_nrm = r_nrm / np.maximum(np.linalg.norm(r_nrm, axis=-1, keepdims=True), 1e-12)
    for precision, tol in (("fp32", 2e-5), ("tc", 5e-3)):
        with torch.no_grad():
            out = m.query_and_decode(pos, tp, exclude=["density", "vertex_offset"], precision=precision)
        assert list(out) == ["features", "perturb_normal"]
        assert out["features"].shape == (3001, 3) and out["perturb_normal"].shape == (3001, 3)
        assert np.abs(out["features"].cpu().numpy() - r_feat).max() < tol
        assert np.abs(out["perturb_normal"].cpu().numpy() - r_nrm).max() < 4 * tol  # unit vectors of small raw outputs
    # a 1-output head through the same per-head path, and include=
    with torch.no_grad():
        d = m.query_and_decode(pos, tp, include=["density"], precision="tc")["density"]
    r_d = fo.material_mlp_head(feats, [sd[f"heads.density.{2 * i}.weight"] for i in range(3)], [sd[f"heads.density.{2 * i}.bias"] for i in range(3)], -1.0, "exp")
    assert np.abs(d.cpu().numpy() / r_d - 1).max() < 5e-3
    with pytest.raises(ValueError):
        m.query_and_decode(pos, tp, include=["density"], exclude=["features"])


def test_lattice_tetgrid_query_vs_reference_and_fp32(golden, tets_dir):
    """The table-based lattice kernel (csrc/tetgrid_tc.cu: layer 0 = C[a][b] + T1[a][c] + T2[b][c] in fp32, hidden layer on
    tcgen05) against the unmodified reference's density / vertex_offset at the grid vertices (golden) and against the fp32
    CUDA-core kernel; then on a grid whose vertex order runs x fastest (axes permuted), against the fp32 kernel."""
    from sculptmate_b200 import runtime
    from sculptmate_b200.sf3d.models.isosurface import detect_lattice
    from sculptmate_b200.tsr.utils import scale_tensor

    g = golden("sf3d_path.npz")
    n = int(g["n"])
    m = _model(g, tets_dir, n, float(g["threshold"]), precision="tc")
    dev = torch.device("cuda")
    assert m.isosurface_helper.lattice is not None and m.isosurface_helper.lattice[0] == (n + 1, n + 1, n + 1)
    tp = torch.from_numpy(g["triplane"]).cuda()
    planes = runtime.prepare_planes_cl(tp)
    packs = [runtime.get_sf3d_head_decoder_pack(m.decoder, "density", dev), runtime.get_sf3d_head_decoder_pack(m.decoder, "vertex_offset", dev)]
    dens, off = runtime.query_tetgrid_tc(planes, packs, (1, 3), (True, False), (-1.0, 0.0), m._lattice_axis_u(dev), m.isosurface_helper.lattice[1])
    torch.cuda.synchronize()
    # the reference's level at every grid vertex (density - threshold, system.py:155) and the deformed grid it produced
    ref_density = g["grid_level"].reshape(-1) + np.float32(g["threshold"])
    e_d = np.abs(dens.cpu().numpy().ravel() / ref_density - 1).max()
    ref_off = np.arctanh(np.clip((g["grid_vertices"] - kuhn_tet_grid(n)[0]) * n, -0.999999, 0.999999))  # isosurface.py:106-113 inverted
    e_o = np.abs(np.tanh(off.cpu().numpy()) - np.tanh(ref_off)).max()
    print(f"lattice tet-grid kernel vs reference: density rel {e_d:.2e}, tanh(vertex_offset) abs {e_o:.2e}")
    assert e_d < 1e-2 and e_o < 1e-2  # fp16 operands in the hidden layer only (the points kernel: 2e-2)
    pos = m._positions(dev)
    r = runtime.sf3d_query(planes, runtime.get_sf3d_heads(m.decoder, dev), -1.0, RADIUS, positions=pos, want=("density_act", "vertex_offset"))
    e_d32 = np.abs(dens.cpu().numpy().ravel() / r["density_act"].cpu().numpy().ravel() - 1).max()
    e_o32 = np.abs(off.cpu().numpy() - r["vertex_offset"].cpu().numpy()).max()
    print(f"lattice tet-grid kernel vs fp32 kernel: density rel {e_d32:.2e}, vertex_offset abs {e_o32:.2e}")
    assert e_d32 < 1e-2 and e_o32 < 1e-2

    # a lattice stored with x running fastest and different extents per axis
    ax = [np.linspace(0, 1, k, dtype=np.float32) for k in (7, 12, 9)]  # x, y, z coordinate lists
    zz, yy, xx = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")  # index order: z slow, y mid, x fast
    verts = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=-1).astype(np.float32)
    lat = detect_lattice(verts)
    assert lat is not None and lat[0] == (9, 12, 7) and lat[1] == (2, 1, 0)
    bbox = m.bbox.to(dev)
    axis_u = []
    for k in range(3):
        c = torch.from_numpy(lat[2][k]).to(dev)
        p_ = scale_tensor(c, (0, 1), (bbox[0, lat[1][k]], bbox[1, lat[1][k]]))
        axis_u.append(scale_tensor(p_, (-RADIUS, RADIUS), (-1, 1)))
    dens2, off2 = runtime.query_tetgrid_tc(planes, packs, (1, 3), (True, False), (-1.0, 0.0), axis_u, lat[1])
    pos2 = scale_tensor(torch.from_numpy(verts).to(dev), (0, 1), bbox).contiguous()
    r2 = runtime.sf3d_query(planes, runtime.get_sf3d_heads(m.decoder, dev), -1.0, RADIUS, positions=pos2, want=("density_act", "vertex_offset"))
    assert np.abs(dens2.cpu().numpy().ravel() / r2["density_act"].cpu().numpy().ravel() - 1).max() < 1e-2
    assert np.abs(off2.cpu().numpy() - r2["vertex_offset"].cpu().numpy()).max() < 1e-2
