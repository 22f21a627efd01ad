#!/usr/bin/env python3
"""Generate the 256-case marching-cubes triangle table used by the CUDA kernels
(``sculptmate_b200/csrc/mc_tables.h``) and by the CPU oracle (``oracle/mc_tables_oracle.h``).

Why generated, not transcribed: the reference calls scikit-image's Lewiner
marching cubes (TripoSR/tsr/models/isosurface.py:46-48); scikit-image and its
33-case LUTs are not available in this environment, so
the in-repo algorithm is a classic 256-case table.  Instead of typing a 4 KB
table from memory, the table is *derived* from a stated rule, and
``tests/test_mc_tables.py`` proves the properties that make it a valid
crack-free table (every crossing edge used, closed oriented loops, face
compatibility between neighbouring cells, <=5 triangles per case).

Conventions (shared by kernel and oracle; also documented in DESIGN.md):
  * array axes (0,1,2) = (x,y,z); a cell at (i,j,k) has corner
    c = 4*di + 2*dj + dk at grid point (i+di, j+dj, k+dk);
  * case index = sum over corners of (value(c) > 0) << c;
  * edge e = 4*a + 2*o1 + o2 is the edge along axis a whose other two offsets
    (in increasing axis order) are (o1,o2); it is *owned* by the grid point
    cell+offset and shared by up to four cells;
  * on a face whose four edges all cross (diagonal corners of equal sign) the
    positive corners are separated (each keeps its own contour segment) -- both
    cells sharing the face see the same four values, hence the same choice, hence
    no cracks;
  * each contour loop is oriented with the positive region on its left when
    seen from outside the cell, so the right-hand normal of every triangle points
    towards increasing value; loops are fan-triangulated from their lowest edge id
    and emitted in order of that id.
"""
from __future__ import annotations

import itertools
import os
import sys
from typing import Dict, List, Tuple

import numpy as np

AXES_OTHER = {0: (1, 2), 1: (0, 2), 2: (0, 1)}


def corner_offsets(c: int) -> Tuple[int, int, int]:
    return (c >> 2) & 1, (c >> 1) & 1, c & 1


def corner_index(d: Tuple[int, int, int]) -> int:
    return 4 * d[0] + 2 * d[1] + d[2]


def edge_endpoints(e: int) -> Tuple[int, int]:
    a, o1, o2 = e >> 2, (e >> 1) & 1, e & 1
    b, c = AXES_OTHER[a]
    d0 = [0, 0, 0]
    d0[b], d0[c] = o1, o2
    d1 = list(d0)
    d1[a] = 1
    return corner_index(tuple(d0)), corner_index(tuple(d1))


def edge_owner(e: int) -> Tuple[Tuple[int, int, int], int]:
    """(offset of the owning grid point relative to the cell, axis)."""
    c0, _ = edge_endpoints(e)
    return corner_offsets(c0), e >> 2


def edge_midpoint(e: int) -> np.ndarray:
    c0, c1 = edge_endpoints(e)
    return (np.array(corner_offsets(c0), float) + np.array(corner_offsets(c1), float)) / 2


def faces() -> List[Tuple[np.ndarray, List[int], List[int]]]:
    """Each face: (outward normal, its 4 corner ids, its 4 edge ids)."""
    out = []
    for axis in range(3):
        for side in (0, 1):
            n = np.zeros(3)
            n[axis] = 1.0 if side else -1.0
            corners = [c for c in range(8) if corner_offsets(c)[axis] == side]
            edges = []
            for e in range(12):
                c0, c1 = edge_endpoints(e)
                if c0 in corners and c1 in corners:
                    edges.append(e)
            assert len(corners) == 4 and len(edges) == 4
            out.append((n, corners, edges))
    return out


FACES = faces()


def case_segments(case: int) -> List[Tuple[int, int]]:
    """Oriented contour segments (edge_from, edge_to) on the six faces of a case."""
    pos = [(case >> c) & 1 for c in range(8)]
    segs: List[Tuple[int, int]] = []
    for n, corners, edges in FACES:
        crossing = [e for e in edges if pos[edge_endpoints(e)[0]] != pos[edge_endpoints(e)[1]]]
        if not crossing:
            continue
        assert len(crossing) in (2, 4)
        pairs: List[Tuple[int, int, int]] = []  # (e1, e2, a positive corner on their side)
        if len(crossing) == 2:
            p = [c for c in corners if pos[c]]
            # any positive corner adjacent to one of the two crossing edges is on the positive side
            pc = next(c for c in p if c in edge_endpoints(crossing[0]))
            pairs.append((crossing[0], crossing[1], pc))
        else:
            # ambiguous face: isolate every positive corner with the two face edges touching it
            for c in corners:
                if pos[c]:
                    es = [e for e in edges if c in edge_endpoints(e)]
                    assert len(es) == 2
                    pairs.append((es[0], es[1], c))
        for e1, e2, pc in pairs:
            m1, m2 = edge_midpoint(e1), edge_midpoint(e2)
            p = np.array(corner_offsets(pc), float)
            s = float(np.dot(n, np.cross(m2 - m1, p - m1)))
            assert abs(s) > 1e-9
            segs.append((e1, e2) if s > 0 else (e2, e1))
    return segs


def case_loops(case: int) -> List[List[int]]:
    segs = case_segments(case)
    nxt: Dict[int, int] = {}
    for a, b in segs:
        assert a not in nxt, "edge has two outgoing segments"
        nxt[a] = b
    assert sorted(nxt.keys()) == sorted(nxt.values()), "segments do not close"
    loops: List[List[int]] = []
    seen = set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop = [start]
        seen.add(start)
        cur = nxt[start]
        while cur != start:
            loop.append(cur)
            seen.add(cur)
            cur = nxt[cur]
        loops.append(loop)  # starts at its lowest edge id because starts are visited ascending
    return loops


def case_triangles(case: int) -> List[Tuple[int, int, int]]:
    tris = []
    for loop in case_loops(case):
        for t in range(1, len(loop) - 1):
            tris.append((loop[0], loop[t], loop[t + 1]))
    return tris


def build_tables():
    tri = np.full((256, 16), -1, dtype=np.int8)
    ntri = np.zeros(256, dtype=np.uint8)
    emask = np.zeros(256, dtype=np.uint16)
    for case in range(256):
        t = case_triangles(case)
        assert len(t) <= 5, (case, len(t))
        ntri[case] = len(t)
        flat = list(itertools.chain.from_iterable(t))
        tri[case, : len(flat)] = flat
        for e in range(12):
            c0, c1 = edge_endpoints(e)
            if ((case >> c0) & 1) != ((case >> c1) & 1):
                emask[case] |= 1 << e
    return tri, ntri, emask


def emit_header(path: str, guard: str, ns_comment: str) -> None:
    tri, ntri, emask = build_tables()
    lines = []
    lines.append("// GENERATED by tools/gen_mc_tables.py -- do not edit by hand.")
    lines.append(f"// {ns_comment}")
    lines.append("// corner c = 4*di+2*dj+dk ; case bit c = (value(c) > 0) ; edge e = 4*axis+2*o1+o2.")
    lines.append("// Triangles' right-hand normals point towards increasing value.")
    lines.append(f"#ifndef {guard}\n#define {guard}")
    lines.append("#ifndef SMB_TABLE_QUAL\n#define SMB_TABLE_QUAL static const\n#endif")
    lines.append("SMB_TABLE_QUAL unsigned char SMB_MC_NTRI[256] = {")
    for r in range(0, 256, 32):
        lines.append("  " + ",".join(str(int(x)) for x in ntri[r : r + 32]) + ",")
    lines.append("};")
    lines.append("SMB_TABLE_QUAL unsigned short SMB_MC_EDGEMASK[256] = {")
    for r in range(0, 256, 16):
        lines.append("  " + ",".join("0x%03x" % int(x) for x in emask[r : r + 16]) + ",")
    lines.append("};")
    lines.append("SMB_TABLE_QUAL signed char SMB_MC_TRI[256][16] = {")
    for c in range(256):
        lines.append("  {" + ",".join("%2d" % int(x) for x in tri[c]) + "},")
    lines.append("};")
    lines.append("// per edge: owner grid-point offset (di,dj,dk) and axis")
    lines.append("SMB_TABLE_QUAL unsigned char SMB_MC_EDGE_OWNER[12][4] = {")
    for e in range(12):
        (di, dj, dk), a = edge_owner(e)
        lines.append(f"  {{{di},{dj},{dk},{a}}},")
    lines.append("};")
    lines.append(f"#endif  // {guard}")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def main() -> None:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    emit_header(
        os.path.join(root, "sculptmate_b200", "csrc", "mc_tables.h"),
        "SMB_MC_TABLES_H",
        "Product copy (CUDA kernels).",
    )
    emit_header(
        os.path.join(root, "oracle", "mc_tables_oracle.h"),
        "SMB_MC_TABLES_ORACLE_H",
        "Oracle copy (test infrastructure).",
    )
    tri, ntri, _ = build_tables()
    print("cases:", 256, "max tris:", int(ntri.max()), "total tris:", int(ntri.sum()))


if __name__ == "__main__":
    sys.exit(main())
