#!/usr/bin/env python3
"""Dev timing of the SF3D mesh path (BASELINE configs[4]): triplane_to_meshes on a synthetic
3x40x384x384 triplane and a Kuhn tet grid of n^3 cubes.  python tools/bench_sf3d.py [n] [iters]"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import baked_triplane  # noqa: E402
from sculptmate_b200 import runtime  # noqa: E402
from sculptmate_b200.sf3d import SF3D, save_tet_grid  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
path = save_tet_grid(os.path.join(tempfile.mkdtemp(), f"tets{n}.npz"), n)
torch.manual_seed(0)
m = SF3D(dict(isosurface_resolution=n, radius=0.87, tets_path=path)).to(dev)
tp = baked_triplane(100, 384, 384).to(dev)
t0 = time.perf_counter()
h = m.isosurface_helper
edges, tets, tet_edges = h.topology(dev)
torch.cuda.synchronize()
print(f"grid n={n}: Nv={h.grid_vertices.shape[0]} Nt={tets.shape[0]} Ne={edges.shape[0]}  topology set-up {time.perf_counter() - t0:.2f} s")
planes = runtime.prepare_planes_cl(tp)
pos = m._positions(dev)
d = runtime.sf3d_query(planes, runtime.get_sf3d_heads(m.decoder, dev), -1.0, 0.87, positions=pos, want=("density_act",))["density_act"]
m.cfg.isosurface_threshold = float(d.median())
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
ts, tq, tm = [], [], []
for i in range(iters + 2):
    a, b, c, e = ev(), ev(), ev(), ev()
    a.record()
    planes = runtime.prepare_planes_cl(tp)
    dec = runtime.sf3d_query(planes, runtime.get_sf3d_heads(m.decoder, dev), -1.0, 0.87, positions=pos, want=("density_act", "vertex_offset"))
    b.record()
    sdf = dec["density_act"] - m.cfg.isosurface_threshold
    c.record()
    mesh = h(sdf.view(-1, 1), dec["vertex_offset"])
    e.record()
    torch.cuda.synchronize()
    if i >= 2:
        ts.append(a.elapsed_time(e)); tq.append(a.elapsed_time(b)); tm.append(c.elapsed_time(e))
print(f"sf3d (fp32 query) triplane_to_mesh n={n}: total {np.median(ts):.3f} ms  (query+heads {np.median(tq):.3f} ms, marching tets {np.median(tm):.3f} ms)  V={mesh.v_pos.shape[0]} F={mesh.t_pos_idx.shape[0]}")
# tensor-core query: fp32 planes, then fp16 planes
tcp = runtime.get_sf3d_points_pack(m.decoder, dev)
for label, planes in (("fp32 planes", runtime.prepare_planes_cl(tp)), ("fp16 planes", runtime.prepare_planes_half(tp))):
  tq = []
  for i in range(iters + 2):
    a, b = ev(), ev()
    a.record()
    r = runtime.query_points_tc(planes, tcp, pos, 0.87, -1.0, align_corners=True, sigmoid_vec=False, want=("out0_act", "vec"))
    b.record()
    torch.cuda.synchronize()
    if i >= 2:
        tq.append(a.elapsed_time(b))
  rel = float((r["out0_act"] / dec["density_act"] - 1).abs().max())
  print(f"sf3d tensor-core query ({label}) n={n}: {np.median(tq):.3f} ms for {pos.shape[0]} points ({pos.shape[0] / np.median(tq) / 1e6:.2f} Gpts/s); density max rel diff vs fp32 {rel:.2e}; "
        f"offset max abs diff {float((r['vec'] - dec['vertex_offset']).abs().max()):.2e}")
