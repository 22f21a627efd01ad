"""CPU: the SF3D oracle (oracle/sf3d_oracle.py, oracle/field_oracle.py) against the fixtures
produced by the unmodified reference (oracle/make_golden_sf3d.py), plus host-side checks of
the sf3d mirror that need no GPU."""
import numpy as np
import pytest
import torch

from oracle import field_oracle as fo
from oracle import sf3d_oracle as so
from sculptmate_b200.sf3d.tets import kuhn_tet_grid


def _sd(g):
    return {k[3:]: g[k] for k in g.files if k.startswith("sd.")}


@pytest.mark.parametrize("name", ["sphere", "torus", "noise"])
def test_marching_tets_oracle_matches_reference(golden, name):
    g = golden("sf3d_mtet.npz")
    n = int(g["n"])
    verts, tets = kuhn_tet_grid(n)
    pos = verts
    if f"{name}_deform" in g.files:
        pos = so.deform_grid(verts, g[f"{name}_deform"], n)
        np.testing.assert_allclose(pos, g[f"{name}_grid"], rtol=0, atol=2e-7)  # tanh differs by an ulp across libms
        pos = g[f"{name}_grid"]
    v, f = so.marching_tets(pos, g[f"{name}_sdf"], tets)
    np.testing.assert_array_equal(f, g[f"{name}_f"])  # vertex numbering and face order bit-exact
    np.testing.assert_array_equal(v, g[f"{name}_v"])  # fp32 operation order restated exactly


def test_static_edge_list_is_the_references_all_edges(golden):
    g = golden("sf3d_mtet.npz")
    _, tets = kuhn_tet_grid(int(g["n"]))
    e = np.sort(tets[:, so.BASE_TET_EDGES].reshape(-1, 2), axis=1)
    np.testing.assert_array_equal(np.unique(e, axis=0), g["all_edges"])


def test_query_and_heads_oracle_match_reference(golden):
    g = golden("sf3d_path.npz")
    sd = _sd(g)
    feats = fo.sf3d_query_triplane(g["positions"], g["triplane"], 0.87)
    assert np.abs(feats - g["features"][0]).max() < 2e-6
    d = fo.material_mlp_head(g["features"][0], *so.heads_from_state_dict(sd, "density"), out_bias=-1.0, activation="trunc_exp")
    o = fo.material_mlp_head(g["features"][0], *so.heads_from_state_dict(sd, "vertex_offset"))
    assert np.abs(d / g["density"][0] - 1).max() < 2e-5
    assert np.abs(o - g["vertex_offset"][0]).max() < 2e-5


def test_triplane_to_mesh_oracle_matches_reference(golden):
    g = golden("sf3d_path.npz")
    n = int(g["n"])
    verts, tets = kuhn_tet_grid(n)
    r = so.triplane_to_mesh(g["triplane"], _sd(g), verts, tets, n, float(g["threshold"]))
    # connectivity depends on the sign of (density - thr): fp32 noise can flip samples that sit
    # within ~1e-6 of the threshold, so compare topology only when the sign pattern agrees
    same_sign = np.array_equal(r["sdf"].reshape(-1) > 0, g["grid_level"].reshape(-1) > 0)
    assert (np.sign(r["sdf"].reshape(-1)) != np.sign(g["grid_level"].reshape(-1))).mean() < 1e-3
    if same_sign:
        np.testing.assert_array_equal(r["t_pos_idx"], g["t_pos_idx"])
        assert np.abs(r["v_pos"] - g["v_pos"]).max() < 5e-4
    # with the reference's own level and deformed grid the mesh is exact
    v, f = so.marching_tets(g["grid_vertices"], g["grid_level"], tets)
    np.testing.assert_array_equal(f, g["t_pos_idx"])
    v = (v * np.float32(2 * 0.87) + np.float32(-0.87)).astype(np.float32)
    np.testing.assert_array_equal(v, g["v_pos"])


def test_material_mlp_state_dict_keys_match_reference(golden):
    from sculptmate_b200.sf3d import MaterialMLP
    from sculptmate_b200.sf3d.system import DEFAULT_DECODER_CFG

    g = golden("sf3d_path.npz")
    m = MaterialMLP(dict(DEFAULT_DECODER_CFG))
    ref_keys = sorted(k[3:] for k in g.files if k.startswith("sd."))
    assert sorted(m.state_dict().keys()) == ref_keys
    m.load_state_dict({k: torch.from_numpy(v) for k, v in _sd(g).items()})
    assert m.cuda_heads_supported()
    # eager CPU forward of a head outside the fused path keeps the reference semantics
    x = torch.from_numpy(g["features"])
    with pytest.raises(ValueError):
        m(x, include=["density"], exclude=["vertex_offset"])


def test_kuhn_grid_is_a_valid_partition():
    v, t = kuhn_tet_grid(4)
    assert v.shape == (125, 3) and t.shape == (6 * 64, 4) and v.min() == 0 and v.max() == 1
    p = v[t].astype(np.float64)
    vol = np.abs(np.einsum("ij,ij->i", p[:, 1] - p[:, 0], np.cross(p[:, 2] - p[:, 0], p[:, 3] - p[:, 0]))) / 6
    assert np.allclose(vol.sum(), 1.0) and np.allclose(vol, vol[0])  # 6n^3 congruent tets tile the unit cube


def test_static_topology_of_the_mirror_helper(tmp_path, golden):
    """MarchingTetrahedraHelper.topology (host-side, torch ops): the int32 edge list is the
    reference's all_edges and tet_edges indexes each tet's six base edges into it."""
    from sculptmate_b200.sf3d import MarchingTetrahedraHelper, save_tet_grid

    g = golden("sf3d_mtet.npz")
    n = int(g["n"])
    h = MarchingTetrahedraHelper(n, save_tet_grid(str(tmp_path / "t.npz"), n))
    np.testing.assert_array_equal(h.all_edges.numpy(), g["all_edges"])
    edges, tets, tet_edges = h.topology(torch.device("cpu"))
    assert edges.dtype == tets.dtype == tet_edges.dtype == torch.int32 and tet_edges.shape == (tets.shape[0], 6)
    e, t, te = edges.numpy(), tets.numpy(), tet_edges.numpy()
    pairs = np.sort(t[:, so.BASE_TET_EDGES].reshape(-1, 2), axis=1)
    np.testing.assert_array_equal(e[te.reshape(-1)], pairs)
    assert h.normalize_grid_deformation(torch.zeros(3, 3)).abs().max() == 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        h(torch.zeros(h.grid_vertices.shape[0], 1), None)  # forward needs the CUDA library


def test_lattice_table_decomposition_equals_the_first_linear(golden):
    """The identity csrc/tetgrid_tc.cu is built on, checked on the CPU against the reference-pinned oracle: at the vertices
    of a lattice-ordered grid (sf3d/models/isosurface.py detect_lattice) the first Linear of a head on the concatenated
    plane features (sf3d/system.py:170-198, network.py:158-178) is the sum of three TWO-index tables,
        W0 . [f_xy; f_xz; f_yz] + b0 = C[a][b] + T1[a][c] + T2[b][c],
    each built from n^2 bilinear interpolations of ONE plane.  Axes deliberately permuted and unevenly sized."""
    from sculptmate_b200.sf3d.models.isosurface import detect_lattice

    g = golden("sf3d_path.npz")
    sd = _sd(g)
    W, b = so.heads_from_state_dict(sd, "density")
    W0, b0 = W[0].astype(np.float64), b[0].astype(np.float64)
    tp = g["triplane"].astype(np.float32)
    Cp = tp.shape[1]
    ax = [np.linspace(0.02, 0.97, k).astype(np.float32) for k in (5, 7, 6)]  # x, y, z coordinate lists in grid units
    zz, xx, yy = np.meshgrid(ax[2], ax[0], ax[1], indexing="ij")  # vertex order: z slow, x mid, y fast
    verts = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], -1).astype(np.float32)
    ext, sdim, coords = detect_lattice(verts)
    assert ext == (6, 5, 7) and sdim == (2, 0, 1)
    r = np.float32(0.87)
    pos = (verts * (r - (-r)) + (-r)).astype(np.float32)  # scale_tensor(grid, (0,1), bbox)
    feats = fo.sf3d_query_triplane(pos, tp, 0.87).astype(np.float64)
    direct = feats @ W0.T + b0  # (N, 64)

    def table(lp, lq):
        """Entries (ip, iq) of the table of lattice indices (lp, lq): the ONE plane both spatial axes belong to."""
        dp, dq = sdim[lp], sdim[lq]
        plane = dp + dq - 1  # {x,y} -> 0, {x,z} -> 1, {y,z} -> 2
        P = np.zeros((ext[lp] * ext[lq], 3), np.float32)
        ip, iq = np.meshgrid(np.arange(ext[lp]), np.arange(ext[lq]), indexing="ij")
        P[:, dp] = coords[lp][ip.ravel()]
        P[:, dq] = coords[lq][iq.ravel()]
        P = (P * (r - (-r)) + (-r)).astype(np.float32)  # the third coordinate is irrelevant for this plane
        f = fo.sf3d_query_triplane(P, tp, 0.87).astype(np.float64)[:, plane * Cp : (plane + 1) * Cp]
        return (f @ W0[:, plane * Cp : (plane + 1) * Cp].T).reshape(ext[lp], ext[lq], -1)

    C, T1, T2 = table(0, 1) + b0, table(0, 2), table(1, 2)
    total = (C[:, :, None, :] + T1[:, None, :, :] + T2[None, :, :, :]).reshape(-1, W0.shape[0])
    assert np.abs(total - direct).max() < 1e-6
