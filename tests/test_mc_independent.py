"""Independent check of the marching-cubes tables the kernels and the oracle ship.

Everything here is written WITHOUT importing tools/gen_mc_tables.py (the generator both table headers come from): the
committed headers are parsed as text, the cube geometry is restated from the convention comment at the top of the header
(corner c = 4*di + 2*dj + dk, case bit c = value(c) > 0, edge e = 4*axis + 2*o1 + o2), and the expectations are
typed in from the literature:

  * the 15 base configurations of Lorensen & Cline (1987), as sets of POSITIVE corners, with the triangle count each
    produces when diagonal positive corners on an ambiguous face are kept separate (the rule this table states);
  * orientation: the right-hand normal of every triangle points towards increasing value;
  * per cube face the number of contour segments is fixed by the face's four signs (0 / 1 / 2) and, on a face shared by
    two cells, both cells must draw the same segments (no cracks).

So an error in the generator cannot hide behind "GPU == oracle": both would then disagree with this file.
"""
import itertools
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = ["oracle/mc_tables_oracle.h", "sculptmate_b200/csrc/mc_tables.h"]


def parse_header(rel):
    txt = open(os.path.join(ROOT, rel)).read()

    def block(name):
        m = re.search(name + r"\[256\](?:\[16\])? = \{(.*?)\n\};", txt, re.S)
        assert m, name
        return [int(x, 0) for x in re.findall(r"-?(?:0x[0-9a-fA-F]+|\d+)", m.group(1))]

    ntri = np.array(block("SMB_MC_NTRI"))
    emask = np.array(block("SMB_MC_EDGEMASK"))
    tri = np.array(block("SMB_MC_TRI")).reshape(256, 16)
    return tri, ntri, emask


# ---- cube geometry restated from the header's convention line (not from the generator)
def corner_xyz(c):
    return np.array([(c >> 2) & 1, (c >> 1) & 1, c & 1])


def edge_ends(e):
    axis, o1, o2 = e >> 2, (e >> 1) & 1, e & 1
    others = [a for a in range(3) if a != axis]
    p = np.zeros(3, int)
    p[others[0]], p[others[1]] = o1, o2
    q = p.copy()
    q[axis] = 1
    return p, q


def corner_id(p):
    return int(4 * p[0] + 2 * p[1] + p[2])


def edge_mid(e):
    p, q = edge_ends(e)
    return (p + q) / 2.0


def case_triangles(tri, ntri, case):
    return [tuple(int(x) for x in tri[case][3 * t : 3 * t + 3]) for t in range(int(ntri[case]))]


def sign_of(case, p):
    return 1 if (case >> corner_id(p)) & 1 else -1


# ---- the 24 proper rotations of the cube as permutations of corner coordinates
def rotations():
    out = []
    for perm in itertools.permutations(range(3)):
        for flips in itertools.product((0, 1), repeat=3):
            m = np.zeros((3, 3), int)
            for r in range(3):
                m[r, perm[r]] = -1 if flips[r] else 1
            if round(np.linalg.det(m)) == 1:
                out.append(m)
    assert len(out) == 24
    return out


def rotate_case(corners, m):
    """corners: iterable of 0/1 xyz triples (positive corners) -> case index after rotating the cube about its centre."""
    case = 0
    for c in corners:
        p = m @ (2 * np.array(c) - 1)
        case |= 1 << corner_id((p + 1) // 2)
    return case


# Lorensen & Cline's 15 configurations: positive corners (x, y, z) and the triangles they produce when positive corners that are
# only diagonal neighbours across a face are NOT joined on that face
BASE_CASES = [
    ((), 0),
    (((0, 0, 0),), 1),
    (((0, 0, 0), (1, 0, 0)), 2),                          # an edge
    (((0, 0, 0), (1, 1, 0)), 2),                          # face diagonal: two separate corners
    (((0, 0, 0), (1, 1, 1)), 2),                          # body diagonal
    (((0, 0, 0), (1, 0, 0), (0, 1, 0)), 3),               # three on a face
    (((0, 0, 0), (1, 0, 0), (1, 1, 1)), 3),               # an edge and the far corner
    (((0, 0, 0), (1, 1, 0), (1, 0, 1)), 3),               # three mutually diagonal corners
    (((0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0)), 2),    # a whole face
    (((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)), 4),    # a corner and its three neighbours: hexagon
    (((0, 0, 0), (0, 0, 1), (1, 1, 0), (1, 1, 1)), 4),    # two opposite edges
    (((0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1)), 4),    # a chain of four
    (((0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 1)), 4),    # three on a face and the far corner
    (((0, 0, 0), (1, 1, 0), (1, 0, 1), (0, 1, 1)), 4),    # four mutually diagonal corners
    (((0, 0, 0), (0, 1, 0), (1, 1, 0), (1, 1, 1)), 4),    # the mirror chain
]


@pytest.mark.parametrize("rel", HEADERS)
def test_lorensen_cline_base_cases_in_all_orientations(rel):
    tri, ntri, emask = parse_header(rel)
    seen = set()
    for corners, expected in BASE_CASES:
        for m in rotations():
            case = rotate_case(corners, m)
            seen.add(case)
            assert ntri[case] == expected, (rel, corners, case, int(ntri[case]), expected)
    # together with the complements (checked structurally below) the 15 configurations reach all 256 cases
    assert len(seen | {255 - c for c in seen}) == 256


@pytest.mark.parametrize("rel", HEADERS)
def test_every_triangle_sits_on_crossing_edges_and_faces_increasing_value(rel):
    tri, ntri, emask = parse_header(rel)
    for case in range(256):
        crossing = set()
        for e in range(12):
            p, q = edge_ends(e)
            if sign_of(case, p) != sign_of(case, q):
                crossing.add(e)
        assert crossing == {e for e in range(12) if (emask[case] >> e) & 1}, (rel, case)
        used = set()
        flux = 0.0
        for a, b, c in case_triangles(tri, ntri, case):
            assert {a, b, c} <= crossing and len({a, b, c}) == 3, (rel, case)
            used |= {a, b, c}
            n = np.cross(edge_mid(b) - edge_mid(a), edge_mid(c) - edge_mid(a))
            assert np.linalg.norm(n) > 1e-9, (rel, case)
            dots = []
            for e in (a, b, c):
                p, q = edge_ends(e)
                d = (q - p) if sign_of(case, q) > 0 else (p - q)  # from the negative to the positive end
                dots.append(float(np.dot(n, d)))
            # fan triangles of a non-planar polygon may tilt against one of their three edges (or, in the 5- and 6-corner
            # cases, stand edge-on to all of them): no triangle may point from positive to negative on balance, and the
            # cell's patch as a whole must point from its negative to its positive corners
            assert sum(dots) > -1e-9, (rel, case, dots)
            flux += sum(dots)
        assert flux > 1e-9 or not crossing, (rel, case, flux)
        assert used == crossing, (rel, case)
        assert (tri[case][3 * int(ntri[case]) :] == -1).all()


def face_list():
    """(axis, side) -> its 4 corners (xyz) and its 4 edges."""
    faces = {}
    for axis in range(3):
        for side in (0, 1):
            corners = [corner_xyz(c) for c in range(8) if corner_xyz(c)[axis] == side]
            edges = [e for e in range(12) if edge_ends(e)[0][axis] == side and edge_ends(e)[1][axis] == side]
            assert len(corners) == 4 and len(edges) == 4
            faces[(axis, side)] = (corners, edges)
    return faces


def boundary_segments(tri, ntri, case):
    """Undirected triangle edges that are not shared by two triangles of the cell = the contour drawn on the cube's faces."""
    cnt = {}
    for a, b, c in case_triangles(tri, ntri, case):
        for u, v in ((a, b), (b, c), (c, a)):
            cnt[(u, v)] = cnt.get((u, v), 0) + 1
    segs = []
    for (u, v), k in cnt.items():
        assert k == 1, "a directed edge used twice inside one cell"
        if (v, u) not in cnt:
            segs.append(frozenset((u, v)))
    return segs


@pytest.mark.parametrize("rel", HEADERS)
def test_face_contours_match_the_face_signs_and_the_neighbouring_cell(rel):
    tri, ntri, emask = parse_header(rel)
    faces = face_list()
    per_face = {}
    for case in range(256):
        segs = boundary_segments(tri, ntri, case)
        for s in segs:  # every boundary segment lies in exactly one face of the cube ...
            homes = [f for f, (_, edges) in faces.items() if s <= set(edges)]
            assert len(homes) == 1, (rel, case, s)
        for f, (corners, edges) in faces.items():
            on_face = sorted(tuple(sorted(s)) for s in segs if s <= set(edges))
            signs = tuple(sign_of(case, p) for p in corners)
            npos = sum(1 for x in signs if x > 0)
            diagonal = npos == 2 and signs[0] == signs[3]  # corners are listed (0,0),(0,1),(1,0),(1,1): equal diagonal = ambiguous
            assert len(on_face) == (0 if npos in (0, 4) else 2 if diagonal else 1), (rel, case, f)
            # ... and what is drawn depends only on the face's own four signs (not on the rest of the cell)
            key = (f, signs)
            assert per_face.setdefault(key, on_face) == on_face, (rel, case, f)
    # a face shared by two cells: (axis, 1) of one is (axis, 0) of the other with the same four values
    for axis in range(3):
        c_hi, e_hi = faces[(axis, 1)]
        c_lo, e_lo = faces[(axis, 0)]

        def across(e):  # the same physical edge named in the neighbour's frame
            p, q = edge_ends(e)
            p, q = p.copy(), q.copy()
            p[axis] = q[axis] = 0
            return next(e2 for e2 in e_lo if (edge_ends(e2)[0] == p).all() and (edge_ends(e2)[1] == q).all())

        for signs in itertools.product((-1, 1), repeat=4):
            hi = per_face[((axis, 1), signs)]
            lo = per_face[((axis, 0), signs)]
            assert sorted(tuple(sorted(across(e) for e in s)) for s in hi) == lo, (rel, axis, signs)


def test_product_and_oracle_tables_are_the_same_numbers():
    a, b = parse_header(HEADERS[0]), parse_header(HEADERS[1])
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
