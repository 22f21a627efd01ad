"""Pin the CPU oracle to the reference: every function in oracle/field_oracle.py and the
wrapper semantics of oracle/mc_oracle.py are checked against fixtures produced by the
UNMODIFIED reference (oracle/make_golden.py)."""
import hashlib

import numpy as np
import torch

from conftest import golden_decoder
from oracle import field_oracle as fo
from oracle import mc_oracle

RADIUS = 0.87
# fp32-vs-fp32 reimplementation noise floor measured against the reference is 1.2e-5 on the
# logits; allow 4x.
ATOL_LOGIT = 5e-5


def test_lattice_axis_and_vertices(golden):
    g = golden("lattice.npz")
    for R in (2, 5, 16, 33, 64):
        ax = fo.grid_axis(R)
        # aten's CPU linspace: one fused multiply-add per element (restated exactly since round 2)
        np.testing.assert_array_equal(ax, g[f"axis_{R}"])
        a1 = fo.scale_tensor(g[f"axis_{R}"], (0, 1), (-RADIUS, RADIUS))
        a2 = fo.scale_tensor(a1, (-RADIUS, RADIUS), (-1, 1))
        # given the same linspace values, the two remaps are bit-exact restatements
        np.testing.assert_array_equal(a1, g[f"axis_scaled_{R}"])
        np.testing.assert_array_equal(a2, g[f"axis_unit_{R}"])
    for R in (2, 5, 16):
        v = fo.grid_vertices(R)
        assert v.shape == (R**3, 3)
        np.testing.assert_array_equal(v, g[f"verts_{R}"])
        # row (i*R+j)*R+k = (x_i, y_j, z_k): x slowest, z fastest
        ref = g[f"verts_{R}"].reshape(R, R, R, 3)
        assert (ref[1, 0, 0] - ref[0, 0, 0])[0] > 0 and (ref[0, 0, 1] - ref[0, 0, 0])[2] > 0


def _check_field(g, triplane):
    ws, bs = golden_decoder(g)
    out = fo.query_triplane(g["positions"], triplane, ws, bs, radius=RADIUS)
    for k in ("density", "features"):
        assert out[k].shape == g[k].shape
        assert np.abs(out[k] - g[k]).max() < ATOL_LOGIT, k
    assert np.abs(out["density_act"] / g["density_act"] - 1).max() < 1e-4
    assert np.abs(out["color"] - g["color"]).max() < 2e-5


def test_query_triplane_small_planes_incl_borders(golden):
    g = golden("field_small.npz")
    _check_field(g, g["triplane"])


def test_query_triplane_64(golden):
    g = golden("field_64.npz")
    torch.manual_seed(int(g["triplane_seed"]))
    tp = torch.randn(3, 40, 64, 64).numpy()
    assert hashlib.sha256(tp.tobytes()).digest() == bytes(g["triplane_sha256"].tobytes()), "seeded triplane differs on this host"
    _check_field(g, tp)


def test_helper_forward_wrapper_semantics(golden):
    """The golden was produced by the reference's MarchingCubeHelper.forward with the oracle MC
    standing in for scikit-image: this pins sign convention, [1,0,2] flip, /(R-1), dtypes."""
    g = golden("helper_sphere.npz")
    R = int(g["resolution"])
    v, f = mc_oracle.helper_forward(g["level_in"], R)
    assert str(g["v_dtype"]) == "torch.float32" and str(g["t_dtype"]) == "torch.int64"
    np.testing.assert_array_equal(v, g["v_pos"])
    np.testing.assert_array_equal(f, g["t_pos_idx"])
    assert v.min() >= 0 and v.max() <= 1


def test_extract_mesh_body(golden):
    """Reference TSR.extract_mesh body (system.py:171-200) vs oracle: density within fp32 noise;
    mesh bit-exact when both sides mesh the same (golden) density grid."""
    g = golden("extract_mesh.npz")
    ws, bs = golden_decoder(g)
    R, thr = int(g["resolution"]), float(g["threshold"])
    dens = fo.grid_density(R, g["triplane"], ws, bs, radius=RADIUS)
    assert np.abs(dens / g["density_act"] - 1).max() < 1e-4
    level_in = -(g["density_act"].reshape(-1, 1) - np.float32(thr))  # system.py:184
    v, f = mc_oracle.helper_forward(level_in, R)
    v = fo.scale_tensor(v, (0, 1), (-RADIUS, RADIUS))  # system.py:185-189
    np.testing.assert_array_equal(v, g["verts"])
    np.testing.assert_array_equal(f, g["faces"])
    # the fused form (val = density - thr, flags) must equal the two-step wrapper form
    v2, f2, _ = mc_oracle.marching_cubes_slab(
        g["density_act"], sub=np.float32(thr), sign=1.0, flags=mc_oracle.FLIP | mc_oracle.DIV | mc_oracle.AFFINE,
        vdiv=float(R - 1.0), vmul=float(RADIUS - (-RADIUS)), vadd=float(-RADIUS),
    )
    np.testing.assert_array_equal(v2, g["verts"])
    np.testing.assert_array_equal(f2, g["faces"])
    # colour query at the vertices (system.py:191-198)
    col = fo.query_triplane(g["verts"], g["triplane"], ws, bs, radius=RADIUS)["color"]
    assert np.abs(col - g["colors"]).max() < 2e-5
