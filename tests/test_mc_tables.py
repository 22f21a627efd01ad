"""The generated 256-case table is a valid crack-free marching-cubes table."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_mc_tables as g  # noqa: E402


def test_every_crossing_edge_used_exactly_once_per_loop_set():
    tri, ntri, emask = g.build_tables()
    for case in range(256):
        loops = g.case_loops(case)
        used = sorted(e for loop in loops for e in loop)
        crossing = [e for e in range(12) if (emask[case] >> e) & 1]
        assert used == crossing
        assert sum(len(l) - 2 for l in loops) == ntri[case] <= 5
        assert all(len(l) >= 3 for l in loops)


def test_triangles_only_use_crossing_edges_and_terminate():
    tri, ntri, emask = g.build_tables()
    for case in range(256):
        row = tri[case]
        n = int(ntri[case])
        assert (row[: 3 * n] >= 0).all() and (row[3 * n :] == -1).all()
        for e in row[: 3 * n]:
            assert (emask[case] >> int(e)) & 1


def test_face_segments_agree_between_neighbouring_cells():
    """Two cells sharing a face see the same 4 values; the contour segments each draws on
    that face must coincide (as undirected edge pairs) -- the no-crack condition."""
    # a face is (axis, side); the neighbour across +axis sees it as (axis, 0)
    for axis in range(3):
        f_hi = next(f for f in g.FACES if f[0][axis] == 1.0)
        f_lo = next(f for f in g.FACES if f[0][axis] == -1.0)

        def face_segments(case, face):
            edges = set(face[2])
            return {frozenset(s) for s in g.case_segments(case) if s[0] in edges and s[1] in edges}

        def to_lo(e):  # same physical edge seen from the neighbour: offset along `axis` 1 -> 0
            c0, c1 = g.edge_endpoints(e)
            d0, d1 = list(g.corner_offsets(c0)), list(g.corner_offsets(c1))
            d0[axis] -= 1
            d1[axis] -= 1
            for e2 in range(12):
                if g.edge_endpoints(e2) == (g.corner_index(tuple(d0)), g.corner_index(tuple(d1))):
                    return e2
            raise AssertionError

        for bits in range(16):  # signs of the 4 shared corners
            hi_corners = f_hi[1]
            lo_corners = f_lo[1]
            for rest_a in (0, 0b1111):  # the far corners of each cell: all negative / all positive
                for rest_b in (0, 0b1111):
                    case_a = 0
                    case_b = 0
                    for idx, c in enumerate(hi_corners):
                        if (bits >> idx) & 1:
                            case_a |= 1 << c
                    for idx, c in enumerate(lo_corners):
                        if (rest_a >> idx) & 1:
                            case_a |= 1 << c
                    for idx, c in enumerate(hi_corners):  # same physical corners, seen at side 0 of cell b
                        d = list(g.corner_offsets(c))
                        d[axis] = 0
                        if (bits >> idx) & 1:
                            case_b |= 1 << g.corner_index(tuple(d))
                        d[axis] = 1
                        if (rest_b >> idx) & 1:
                            case_b |= 1 << g.corner_index(tuple(d))
                    seg_a = {frozenset(to_lo(e) for e in s) for s in face_segments(case_a, f_hi)}
                    seg_b = face_segments(case_b, f_lo)
                    assert seg_a == seg_b, (axis, bits, case_a, case_b)


def test_orientation_normals_point_to_positive():
    """Single positive corner: the triangle's right-hand normal points towards it."""
    for c in range(8):
        (a, b, d), = g.case_triangles(1 << c)
        pa, pb, pd = g.edge_midpoint(a), g.edge_midpoint(b), g.edge_midpoint(d)
        n = np.cross(pb - pa, pd - pa)
        corner = np.array(g.corner_offsets(c), float)
        assert np.dot(n, corner - pa) > 0


def test_committed_headers_match_generator(tmp_path):
    for rel, guard in (("sculptmate_b200/csrc/mc_tables.h", "SMB_MC_TABLES_H"), ("oracle/mc_tables_oracle.h", "SMB_MC_TABLES_ORACLE_H")):
        out = tmp_path / "t.h"
        note = "Product copy (CUDA kernels)." if "csrc" in rel else "Oracle copy (test infrastructure)."
        g.emit_header(str(out), guard, note)
        assert out.read_text() == open(os.path.join(ROOT, rel)).read(), rel
