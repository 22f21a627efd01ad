"""Host logic of the SF3D lattice path: detection of lattice-ordered tet grids (sculptmate_b200/sf3d/models/isosurface.py)."""
import numpy as np

from sculptmate_b200.sf3d.models.isosurface import detect_lattice
from sculptmate_b200.sf3d.tets import kuhn_tet_grid


def test_kuhn_grid_is_a_lattice_in_x_major_order():
    v, _ = kuhn_tet_grid(6)
    ext, sdim, coords = detect_lattice(v)
    assert ext == (7, 7, 7) and sdim == (0, 1, 2)
    for c in coords:
        np.testing.assert_array_equal(c, np.linspace(0, 1, 7, dtype=np.float32))


def test_permuted_and_ragged_lattices():
    ax = [np.array([0.0, 0.3, 1.0], np.float32), np.linspace(0, 1, 5, dtype=np.float32), np.array([0.1, 0.2, 0.4, 0.8], np.float32)]
    yy, zz, xx = np.meshgrid(ax[1], ax[2], ax[0], indexing="ij")  # y slow, z mid, x fast; unevenly spaced lists
    v = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], -1).astype(np.float32)
    ext, sdim, coords = detect_lattice(v)
    assert ext == (5, 4, 3) and sdim == (1, 2, 0)
    np.testing.assert_array_equal(coords[0], ax[1])
    np.testing.assert_array_equal(coords[1], ax[2])
    np.testing.assert_array_equal(coords[2], ax[0])


def test_non_lattices_are_rejected():
    v, _ = kuhn_tet_grid(5)
    w = v.copy()
    w[17, 1] += 1e-3  # one displaced vertex
    assert detect_lattice(w) is None
    rng = np.random.RandomState(0)
    assert detect_lattice(rng.rand(216, 3).astype(np.float32)) is None
    assert detect_lattice(v[rng.permutation(len(v))]) is None  # the same points in another order
    # body-centred cubic points (what quartet-style tetrahedralisations use): corners then cell centres
    c = (v.reshape(6, 6, 6, 3)[:-1, :-1, :-1] + 0.1).reshape(-1, 3)
    assert detect_lattice(np.concatenate([v, c]).astype(np.float32)) is None
    assert detect_lattice(v[:7]) is None
