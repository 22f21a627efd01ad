"""Drop-in for the extract_mesh half of /root/reference/TripoSR/tsr/system.py.

``TSR`` here carries only what the hot path needs -- ``decoder`` (NeRFMLP),
``renderer`` (TriplaneNeRFRenderer) and ``isosurface_helper`` -- with the reference's
method names and signatures:

  set_marching_cubes_resolution(resolution)                        system.py:118-124
  extract_mesh(scene_codes, enable_texture=False, mesh_name='NewMesh',
               resolution=256, threshold=25.0) -> None             system.py:171-200
  import_obj_blender(verts, faces, vertex_colors=None, name=...)   system.py:127-168 (sink)

The image->triplane half (tokenizers, backbone, post-processor; system.py:68-115) is
out of scope and stays in the reference; INTEGRATION.md shows how the reference's own
TSR adopts this path (bind ``extract_mesh`` / swap the three sub-modules).

Differences from the reference, all internal:
  * no (R^3,3) position tensor, no H2D of 201 MB per call: lattice coordinates are three
    R-entry tables built with the reference's own torch ops and positions are formed
    in-kernel;
  * density never leaves the GPU: the fused tensor-core kernel writes the (R,R,R) grid,
    the CUDA marching cubes reads it with ``val = density - threshold`` folded in
    (the reference negates twice, system.py:184 + isosurface.py:45);
  * vertices come out already divided by (R-1) and rescaled to (-radius, radius)
    (isosurface.py:53, system.py:185-189), faces already flipped (isosurface.py:52).
"""
from __future__ import annotations

import inspect
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch

from .. import runtime
from .._capi import MC_AFFINE, MC_DIV, MC_FACES_I32, MC_FLIP
from . import blender_io
from .models.isosurface import MarchingCubeHelper
from .models.nerf_renderer import TriplaneNeRFRenderer
from .models.network_utils import NeRFMLP
from .utils import BaseModule, scale_tensor

# TripoSR/checkpoints/config.yaml:25-37
DEFAULT_DECODER_CFG = dict(in_channels=120, n_neurons=64, n_hidden_layers=9, activation="silu")
DEFAULT_RENDERER_CFG = dict(
    radius=0.87, feature_reduction="concat", density_activation="exp", density_bias=-1.0, num_samples_per_ray=128
)


class TSR(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        decoder: dict = field(default_factory=lambda: dict(DEFAULT_DECODER_CFG))
        renderer: dict = field(default_factory=lambda: dict(DEFAULT_RENDERER_CFG))

    cfg: Config

    def configure(self) -> None:
        self.decoder = NeRFMLP(self.cfg.decoder)
        self.renderer = TriplaneNeRFRenderer(self.cfg.renderer)
        self.isosurface_helper: Optional[MarchingCubeHelper] = None
        self._axis_cache = {}
        # where extract_mesh delivers each mesh; the reference calls bpy here
        self.mesh_sink: Optional[Callable] = None
        self.meshes: List[Tuple[np.ndarray, np.ndarray, Optional[np.ndarray], str]] = []
        # index dtype of the faces handed to the sink: int64 like the reference's LongTensor, or int32 (Blender's width)
        self.face_index_dtype = torch.int64

    # ------------------------------------------------------------------ API
    def set_marching_cubes_resolution(self, resolution: int):
        if self.isosurface_helper is not None and self.isosurface_helper.resolution == resolution:
            return
        self.isosurface_helper = MarchingCubeHelper(resolution)

    def import_obj_blender(self, verts, faces, vertex_colors=None, name="NewMesh", **extra):
        """Sink with the reference's signature.  Inside Blender, assign the reference's own
        ``TSR.import_obj_blender`` to ``mesh_sink``; elsewhere meshes are collected in
        ``self.meshes``."""
        if self.mesh_sink is not None:
            return self.mesh_sink(verts, faces, vertex_colors, name=name, **extra)
        self.meshes.append((verts, faces, vertex_colors, name))

    def _axis(self, resolution: int, device: torch.device) -> torch.Tensor:
        key = (resolution, str(device), float(self.renderer.cfg.radius))
        if key not in self._axis_cache:
            self._axis_cache[key] = runtime.lattice_axis(
                resolution, self.renderer.cfg.radius, self.isosurface_helper.points_range, device=device
            )
        return self._axis_cache[key]

    def extract_mesh_tensors(
        self, scene_code: torch.Tensor, resolution: int, threshold: float, precision: str = "tc",
        faces_dtype: torch.dtype = torch.int64,
    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """One scene code -> (v_pos (V,3) fp32 in (-radius,radius), t_pos_idx (F,3) int64), on device.

        ``precision="tc"`` (default): the whole path is ONE call into the C library (``smb_extract_mesh_device``:
        prepare -> tcgen05 lattice kernel that also ballots the cube-case bits -> count -> totals -> emit queued back
        to back), so no Python runs between the launches.  ``"fp32"``: reference-precision CUDA-core lattice kernel +
        the stand-alone classification pass (parity anchor).  ``faces_dtype=torch.int32`` delivers the same indices
        at the width Blender's loop arrays use."""
        runtime._require_cuda(scene_code, "scene_code")
        self.set_marching_cubes_resolution(resolution)
        R = resolution
        radius = self.renderer.cfg.radius
        if precision == "tc":
            self.renderer._check_supported()
            ex = runtime.get_mesh_extractor(
                self.decoder, radius, self.renderer.cfg.density_bias, int(scene_code.shape[-2]), int(scene_code.shape[-1]), scene_code.device
            )
            with torch.no_grad():
                return ex.extract(scene_code, R, float(threshold), faces_dtype=faces_dtype, axis_u=self._axis(R, scene_code.device))
        # helper(-(density - threshold)) then level = -input  ==  val = density - threshold
        with torch.no_grad():
            density = self.renderer.query_lattice(
                self.decoder, scene_code, R, axis_u=self._axis(R, scene_code.device), precision=precision,
            )
        flags = MC_FLIP | MC_DIV | MC_AFFINE | (MC_FACES_I32 if faces_dtype == torch.int32 else 0)
        v_pos, t_pos_idx, pend = runtime.mc_extract(
            density, sub=float(threshold), sign=1.0, flags=flags,
            vdiv=float(R - 1.0), vmul=float(radius - (-radius)), vadd=float(-radius),
        )
        if pend.nverts == 0 or pend.ntris == 0:
            runtime.raise_for_empty_surface(density, float(threshold), 1.0)
        return v_pos, t_pos_idx

    def extract_mesh(self, scene_codes, enable_texture=False, mesh_name="NewMesh", resolution: int = 256, threshold: float = 25.0):
        """system.py:171-200: per scene code, mesh -> (optional colour query at the vertices) -> sink, which receives
        numpy arrays like the reference's ``import_obj_blender`` does.  The device->host copies go through pinned
        memory (torch's caching host allocator), all queued before ONE synchronisation, instead of the reference's
        pageable ``.cpu().numpy()`` per array; ``face_index_dtype = torch.int32`` hands the sink the index width
        Blender stores (a third less PCIe traffic), the default int64 is the reference's LongTensor."""
        for scene_code in scene_codes:
            if not enable_texture:
                # mesh only: the library's slab pipeline moves each slab's part of the mesh to (caller-owned) pinned host
                # memory while the next slab is computed; only the tail of the 58 MB copy (at 256^3) is exposed
                runtime._require_cuda(scene_code, "scene_code")
                self.set_marching_cubes_resolution(resolution)
                self.renderer._check_supported()
                ex = runtime.get_mesh_extractor(self.decoder, self.renderer.cfg.radius, self.renderer.cfg.density_bias,
                                                int(scene_code.shape[-2]), int(scene_code.shape[-1]), scene_code.device)
                with torch.no_grad():
                    hv, hf = ex.extract_to_host(scene_code, resolution, float(threshold), faces_dtype=self.face_index_dtype,
                                                axis_u=self._axis(resolution, scene_code.device))
                self.import_obj_blender(hv.numpy(), hf.numpy(), None, name=mesh_name)
                continue
            v_pos, t_pos_idx = self.extract_mesh_tensors(scene_code, resolution, threshold, faces_dtype=self.face_index_dtype)
            staged = [self._stage(v_pos), self._stage(t_pos_idx)]
            extra_staged = None
            if enable_texture:
                with torch.no_grad():
                    color = self.renderer.query_triplane(self.decoder, v_pos, scene_code, precision="tc")["color"]
                staged.append(self._stage(color))
                if self._sink_takes_loop_colors():
                    # the per-loop RGBA array the reference builds element by element (system.py:133-146),
                    # gathered on the GPU for sinks that foreach_set it (tsr/blender_io.py)
                    extra_staged = self._stage(blender_io.loop_colors(color, t_pos_idx.long()))
            torch.cuda.current_stream(v_pos.device).synchronize()
            extra = {} if extra_staged is None else {"loop_colors": extra_staged.numpy()}
            color_np = staged[2].numpy() if enable_texture else None
            self.import_obj_blender(staged[0].numpy(), staged[1].numpy(), color_np, name=mesh_name, **extra)

    @staticmethod
    def _stage(t: torch.Tensor) -> torch.Tensor:
        """Asynchronous device -> pinned-host copy; the caller synchronises once for all staged arrays."""
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        return h

    def _sink_takes_loop_colors(self) -> bool:
        if self.mesh_sink is None:
            return False
        try:
            return "loop_colors" in inspect.signature(self.mesh_sink).parameters
        except (TypeError, ValueError):
            return False

    # the reference-shaped slow path, kept for parity tests of the wrapper semantics:
    # grid_vertices -> scale_tensor -> query_triplane -> helper(-(density - threshold))
    def extract_mesh_unfused(self, scene_code: torch.Tensor, resolution: int, threshold: float):
        self.set_marching_cubes_resolution(resolution)
        h = self.isosurface_helper
        with torch.no_grad():
            density = self.renderer.query_triplane(
                self.decoder,
                scale_tensor(h.grid_vertices.to(scene_code.device), h.points_range, (-self.renderer.cfg.radius, self.renderer.cfg.radius)),
                scene_code,
            )["density_act"]
        v_pos, t_pos_idx = h(-(density - threshold))
        v_pos = scale_tensor(v_pos, h.points_range, (-self.renderer.cfg.radius, self.renderer.cfg.radius))
        return v_pos, t_pos_idx
