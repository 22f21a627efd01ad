import os, sys, tempfile, torch
sys.path.insert(0, os.getcwd())
from bench import baked_triplane, RADIUS
from sculptmate_b200 import runtime
from sculptmate_b200.sf3d import SF3D, save_tet_grid
n = 64
path = save_tet_grid(os.path.join(tempfile.mkdtemp(), f"tets{n}.npz"), n)
torch.manual_seed(0)
dev = torch.device("cuda:0")
m = SF3D(dict(isosurface_resolution=n, radius=RADIUS, tets_path=path)).to(dev)
tp = baked_triplane(200, 384, 384).to(dev)
h = m.isosurface_helper
planes = runtime.prepare_planes_cl(tp)
packs = [runtime.get_sf3d_head_decoder_pack(m.decoder, k, dev) for k in ("density", "vertex_offset")]
ax = m._lattice_axis_u(dev)
d0, _ = runtime.query_tetgrid_tc(planes, packs, (1, 3), (True, False), (-1.0, 0.0), ax, h.lattice[1])
thr = float(d0.median()) * 1.000001234
a = d0 - thr
d1, _ = runtime.query_tetgrid_tc(planes, packs, (1, 3), (True, False), (-1.0, 0.0), ax, h.lattice[1], out_sub=(thr, 0.0))
d2, _ = runtime.query_tetgrid_tc(planes, packs, (1, 3), (True, False), (-1.0, 0.0), ax, h.lattice[1])
print("repeatable:", bool(torch.equal(d0, d2)), " in-kernel == torch:", bool(torch.equal(a, d1)), " mismatches", int((a != d1).sum()), "max abs", float((a - d1).abs().max()),
      " sign flips", int(((a > 0) != (d1 > 0)).sum()))
