import os, sys, tempfile, torch
sys.path.insert(0, os.getcwd())
from bench import baked_triplane, RADIUS
from sculptmate_b200 import runtime
from sculptmate_b200.sf3d import SF3D, save_tet_grid
n = 160
path = save_tet_grid(os.path.join(tempfile.mkdtemp(), f"tets{n}.npz"), n)
torch.manual_seed(0)
dev = torch.device("cuda:0")
m = SF3D(dict(isosurface_resolution=n, radius=RADIUS, tets_path=path)).to(dev)
tp = baked_triplane(200, 384, 384).to(dev)
h = m.isosurface_helper
h.topology(dev)
pos = m._positions(dev)
d = runtime.sf3d_query(runtime.prepare_planes_cl(tp), runtime.get_sf3d_heads(m.decoder, dev), -1.0, RADIUS, positions=pos, want=("density_act",))["density_act"]
m.cfg.isosurface_threshold = float(d.median())
for i in range(3):
    mesh = m.triplane_to_meshes(tp[None])[0]
torch.cuda.synchronize()
torch.cuda.profiler.start()
mesh = m.triplane_to_meshes(tp[None])[0]
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(mesh.v_pos.shape, mesh.t_pos_idx.shape)
