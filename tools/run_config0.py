#!/usr/bin/env python3
"""BASELINE configs[0]: the reference AS IS on the CPU -- TSR (random-init, torch.manual_seed(0)) on a synthetic 512x512
image -> scene codes -> TSR.extract_mesh(resolution=128) -- run through oracle/ref_shim.py where /root/reference exists
(the build container; the GPU boxes do not have it).  The only substitutions: skimage.measure.marching_cubes (not installed
anywhere here) is the in-repo C oracle MC, bpy's sink is captured, and the threshold is the median density (random-init
density never reaches the default 25.0: SURVEY fact 2).  Prints one JSON line; the committed copy is
profiles/r02g_config0_reference_cpu.json.  Then the same scene code goes through this repository's CPU port and
(if a GPU is present) nothing else: this tool is baseline-only."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mc_oracle, ref_shim  # noqa: E402


def main() -> int:
    if not ref_shim.reference_available():
        print(json.dumps({"unavailable": "/root/reference is not present on this machine"}))
        return 0
    ref_shim.load_triposr()
    import skimage.measure as skm
    import tsr.system as ref_system
    from omegaconf import OmegaConf

    def mc(level, iso):  # isosurface.py:46-48 with the in-repo oracle (index units, array-axis order, no flip: the reference flips itself)
        v, f, _ = mc_oracle.marching_cubes_slab(np.ascontiguousarray(level, dtype=np.float32), sub=np.float32(iso), flags=0)
        return v, f.astype(np.int32), None, None

    skm.marching_cubes = mc
    import tsr.models.isosurface as iso_mod

    iso_mod.measure.marching_cubes = mc
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    cfg = OmegaConf.load(os.path.join(ref_shim.REFERENCE_ROOT, "TripoSR", "checkpoints", "config.yaml"))
    OmegaConf.resolve(cfg)
    t0 = time.perf_counter()
    model = ref_system.TSR(cfg)
    model.renderer.set_chunk_size(8192)  # generate.py:11,25
    t_build = time.perf_counter() - t0
    captured = []
    model.import_obj_blender = lambda verts, faces, vertex_colors=None, name="NewMesh": captured.append((verts, faces))
    img = (np.random.RandomState(0).rand(512, 512, 3) * 255).astype(np.uint8)
    from PIL import Image

    with torch.no_grad():
        t0 = time.perf_counter()
        scene_codes = model([Image.fromarray(img)], "cpu")
        t_fwd = time.perf_counter() - t0
        R = 128
        h = iso_mod.MarchingCubeHelper(32)
        pos = ref_system.scale_tensor(h.grid_vertices, h.points_range, (-model.renderer.cfg.radius, model.renderer.cfg.radius))
        thr = float(model.renderer.query_triplane(model.decoder, pos, scene_codes[0])["density_act"].median())
        t0 = time.perf_counter()
        model.extract_mesh(scene_codes, resolution=R, threshold=thr)
        t_ext = time.perf_counter() - t0
    v, f = captured[0]
    print(json.dumps({
        "config": "BASELINE configs[0]: reference TSR.forward (512x512 synthetic image) -> TSR.extract_mesh(resolution=128) on CPU, random-init weights, unmodified reference code via oracle/ref_shim.py",
        "cpu": {"threads": torch.get_num_threads(), "where": "build container (8 vCPU Xeon), not the GPU box"},
        "scene_codes_shape": list(scene_codes.shape), "model_build_s": t_build, "forward_s": t_fwd,
        "extract_mesh_s": t_ext, "extract_mesh_points_per_s": R**3 / t_ext, "threshold": thr,
        "mesh": {"verts": int(len(v)), "tris": int(len(f)), "verts_dtype": str(v.dtype), "faces_dtype": str(f.dtype)},
        "substitutions": "skimage.measure.marching_cubes -> oracle/mc_oracle.c (skimage is not installed); bpy sink captured; threshold = median density",
    }))
    return 0


if __name__ == "__main__":
    sys.exit(main())
