"""Properties of the in-repo marching cubes (the oracle that defines 'bit-exact')."""
import numpy as np
import pytest

from conftest import mesh_topology, volume
from oracle import mc_oracle


@pytest.mark.parametrize("kind,chi", [("sphere", 2), ("torus", 0)])
def test_closed_manifold_euler_volume(kind, chi):
    R = 40
    g = volume(kind, R)
    v, f, c = mc_oracle.marching_cubes_slab(g, flags=mc_oracle.FLIP)
    closed, x, vol = mesh_topology(v, f)
    assert closed and x == chi
    assert vol > 0  # after the reference's [1,0,2] flip normals point outwards
    if kind == "sphere":
        h = 2.0 / (R - 1)
        assert abs(vol * h**3 / (4 / 3 * np.pi * 0.6**3) - 1) < 0.02
    assert c.nverts == len(v) and c.ntris == len(f)
    assert c.npos == int((g > 0).sum()) and c.nneg == int((g < 0).sum())


def test_noise_is_watertight_away_from_boundary():
    """Random signs hit every one of the 256 cases incl. all ambiguous faces: no cracks."""
    R = 20
    g = volume("noise", R, seed=3)
    g[0, :, :] = g[-1, :, :] = g[:, 0, :] = g[:, -1, :] = g[:, :, 0] = g[:, :, -1] = -1.0
    cases = mc_oracle.cube_cases(g)
    assert len(np.unique(cases)) == 256
    v, f, _ = mc_oracle.marching_cubes_slab(g)
    closed, _, _ = mesh_topology(v, f, strict=False)
    assert closed


def test_vertices_lie_on_edges_and_interpolate():
    g = volume("smooth", 18, seed=1)
    v, f, _ = mc_oracle.marching_cubes_slab(g)
    frac = v - np.floor(v)
    assert ((frac > 0).sum(axis=1) <= 1).all()  # at most one non-integer coordinate
    assert f.min() == 0 and f.max() == len(v) - 1
    assert len(np.unique(f)) == len(v)  # no orphan vertices, no duplicates by construction


@pytest.mark.parametrize("kind", ["sphere", "gyroid", "noise"])
@pytest.mark.parametrize("world", [2, 3, 5])
def test_slabs_concatenate_bit_exactly(kind, world):
    from sculptmate_b200.dist import slab_partition

    R = 21
    g = volume(kind, R, seed=2)
    kw = dict(sub=0.01, sign=1.0, flags=mc_oracle.FLIP | mc_oracle.DIV, vdiv=float(R - 1))
    v_ref, f_ref, _ = mc_oracle.marching_cubes_slab(g, **kw)
    vs, fs, off = [], [], 0
    for r, (a, b) in enumerate(slab_partition(R, world)):
        v, f, c = mc_oracle.marching_cubes_slab(g[a : b + 1], x_origin=a, emit_last_plane=(r == world - 1), **kw)
        vs.append(v)
        fs.append(f + off)
        off += c.nverts
    np.testing.assert_array_equal(np.concatenate(vs), v_ref)
    np.testing.assert_array_equal(np.concatenate(fs), f_ref)


def test_skimage_like_errors():
    R = 8
    with pytest.raises(ValueError):
        mc_oracle.marching_cubes(np.ones((R, R, R), np.float32), 0.0)
    with pytest.raises(ValueError):
        mc_oracle.marching_cubes(-np.ones((R, R, R), np.float32), 0.0)
    z = -np.ones((R, R, R), np.float32)
    z[3, 3, 3] = 0.0  # level inside [min,max] but nothing crosses
    with pytest.raises(RuntimeError):
        mc_oracle.marching_cubes(z, 0.0)


def test_ragged_shapes_and_minimum_size():
    g = volume("smooth", 12, seed=5)[:, :7, :9].copy()
    v, f, c = mc_oracle.marching_cubes_slab(g)
    assert c.nverts == len(v) and (f < len(v)).all()
    v, f, c = mc_oracle.marching_cubes_slab(np.array([[[1.0, -1.0], [-1.0, -1.0]], [[-1.0, -1.0], [-1.0, -1.0]]], np.float32))
    assert len(v) == 3 and len(f) == 1
