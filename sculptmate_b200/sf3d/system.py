"""Drop-in for the mesh-extraction half of /root/reference/StableFast/sf3d/system.py:

  query_triplane(positions, triplanes) -> (B,N,3*Cp)             system.py:170-198
  triplane_to_meshes(triplanes) -> list[Mesh]                    system.py:141-168

``SF3D`` here carries only ``decoder`` (MaterialMLP), ``isosurface_helper`` and ``bbox``;
the image -> triplane half, texture baking and material estimation stay in the reference
(INTEGRATION.md shows how the reference's SF3D binds these two methods).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch

from .. import runtime
from ..tsr.utils import BaseModule, scale_tensor
from .models.isosurface import MarchingTetrahedraHelper
from .models.mesh import Mesh
from .models.network import MaterialMLP
from .models.network import get_activation as network_activation

# StableFast/checkpoints/config.yaml:45-65 (the two heads of the mesh path)
DEFAULT_DECODER_CFG = dict(
    in_channels=120, n_neurons=64, activation="silu",
    heads=[
        dict(name="density", out_channels=1, out_bias=-1.0, n_hidden_layers=2, output_activation="trunc_exp"),
        dict(name="vertex_offset", out_channels=3, n_hidden_layers=2),
    ],
)


class SF3D(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        isosurface_resolution: int = 160
        isosurface_threshold: float = 10.0
        radius: float = 0.87
        tets_path: str = ""
        precision: str = "tc"  # "tc": tcgen05 kernel (fp16 operands); "fp32": CUDA-core kernel at reference precision
        lattice_path: bool = True  # take the table-based kernel when the tet grid's vertices form a lattice (precision "tc")
        decoder: dict = field(default_factory=lambda: dict(DEFAULT_DECODER_CFG))

    cfg: Config

    def configure(self) -> None:
        self.decoder = MaterialMLP(self.cfg.decoder)
        r = self.cfg.radius
        self.register_buffer("bbox", torch.as_tensor([[-r, -r, -r], [r, r, r]], dtype=torch.float32))
        self.isosurface_helper = MarchingTetrahedraHelper(self.cfg.isosurface_resolution, self.cfg.tets_path)
        self._grid_positions = None
        self._lattice_axes = None
        self._affine = None

    # ------------------------------------------------------------------ API
    def query_triplane(self, positions: torch.Tensor, triplanes: torch.Tensor) -> torch.Tensor:
        batched = positions.ndim == 3
        if not batched:
            triplanes = triplanes[None, ...]
            positions = positions[None, ...]
        assert triplanes.ndim == 5 and positions.ndim == 3
        outs = []
        for b in range(triplanes.shape[0]):
            planes = runtime.prepare_planes_cl(triplanes[b])
            res = runtime.sf3d_query(planes, None, 0.0, self.cfg.radius, positions=positions[b], want=("features",))
            outs.append(res["features"])
        return torch.stack(outs, dim=0)  # the reference keeps the batch dim it adds (system.py:175-179,196)

    def query_and_decode(self, positions: torch.Tensor, triplane: torch.Tensor, include=None, exclude=None, precision: str = None):
        """``decoder(query_triplane(positions, triplane)[0], include=..., exclude=...)`` as fused kernels -- the texel-space
        query of the texture bake, sf3d/system.py:375-378 (``exclude=["density", "vertex_offset"]``: heads ``features``,
        ``perturb_normal``).  positions (N,3) in (-radius, radius); returns the same dict of activated head outputs.
        With ``precision="tc"`` (default: ``cfg.precision``) every head with <= 3 outputs runs gather + MLP on the tensor
        cores (fp16 operands, fp32 accumulate), one pass per head; anything else goes through the fp32 feature kernel and
        the head's torch modules, as does ``precision="fp32"``."""
        if include is not None and exclude is not None:
            raise ValueError("Cannot specify both include and exclude.")
        heads = [h for h in self.decoder.cfg.heads if (include is None or h.name in include) and (exclude is None or h.name not in exclude)]
        precision = precision or self.cfg.precision
        dev = triplane.device
        out, rest = {}, []
        if precision == "tc":
            planes = runtime.prepare_planes_half(triplane)
            for h in heads:
                try:
                    pack = runtime.get_sf3d_head_pack(self.decoder, h.name, dev)
                except NotImplementedError:
                    rest.append(h.name)
                    continue
                raw = runtime.query_points_tc(planes, pack, positions, self.cfg.radius, 0.0, align_corners=True, sigmoid_vec=False,
                                              want=("vec",))["vec"][:, : h.out_channels]
                out[h.name] = network_activation(h.output_activation)(raw + h.out_bias)
        else:
            rest = [h.name for h in heads]
        if rest:
            feats = self.query_triplane(positions, triplane)[0]
            out.update(self.decoder(feats, include=rest))
        return {h.name: out[h.name] for h in heads}

    def _positions(self, device: torch.device) -> torch.Tensor:
        if self._grid_positions is None or self._grid_positions.device != device:
            h = self.isosurface_helper
            self._grid_positions = scale_tensor(h.grid_vertices.to(device), h.points_range, self.bbox.to(device)).contiguous()
        return self._grid_positions

    def _v_pos_affine(self, device: torch.device):
        """scale_tensor(v, points_range, bbox) as the 8 floats of ``runtime.marching_tets``: the differences are taken in
        fp32 exactly as utils.py:228-230 takes them (python-float subtraction for the input range, tensor for the bbox)."""
        if self._affine is None:
            h = self.isosurface_helper
            bbox = self.bbox.detach().to("cpu", torch.float32)
            span = bbox[1] - bbox[0]
            lo, hi = h.points_range
            self._affine = [float(lo), float(hi - lo), *[float(x) for x in span], *[float(x) for x in bbox[0]]]
        return self._affine

    def _lattice_axis_u(self, device: torch.device):
        """Per-index coordinate lists of a lattice-ordered grid, taken through the same two ``scale_tensor`` calls as every
        vertex (grid -> bbox :147-151, then (-radius, radius) -> (-1, 1) :175-177): elementwise ops, so the values are
        bit-identical to what the per-vertex path feeds grid_sample."""
        if self._lattice_axes is None or self._lattice_axes[0] != device:
            h = self.isosurface_helper
            _, sdim, coords = h.lattice
            bbox = self.bbox.to(device)
            axes = []
            for k in range(3):
                c = torch.from_numpy(coords[k]).to(device)
                pos = scale_tensor(c, h.points_range, (bbox[0, sdim[k]], bbox[1, sdim[k]]))
                axes.append(scale_tensor(pos, (-self.cfg.radius, self.cfg.radius), (-1, 1)).contiguous())
            self._lattice_axes = (device, axes)
        return self._lattice_axes[1]

    def triplane_to_meshes(self, triplanes: torch.Tensor) -> List[Mesh]:
        meshes = []
        h = self.isosurface_helper
        for i in range(triplanes.shape[0]):
            triplane = triplanes[i]
            dev = triplane.device
            grid_vertices = self._positions(dev)  # scale_tensor(grid, points_range, bbox)   :147-151
            level_done = False
            if self.decoder.cuda_heads_supported():
                tc = self.cfg.precision == "tc"
                lattice = tc and h.lattice is not None and self.cfg.lattice_path
                planes = None if lattice else (runtime.prepare_planes_half(triplane) if tc else runtime.prepare_planes_cl(triplane))
                dens_spec = self.decoder.head_spec("density")
                # query_triplane + decoder(include=[vertex_offset, density]) fused into one kernel  :153-154
                if tc and h.lattice is not None and self.cfg.lattice_path:
                    # lattice-ordered grid: layer 0 from three n^2 tables instead of n^3 plane gathers (csrc/tetgrid_tc.cu)
                    offs = self.decoder.head_spec("vertex_offset")
                    density, deform = runtime.query_tetgrid_tc(
                        runtime.prepare_planes_cl(triplane),
                        [runtime.get_sf3d_head_decoder_pack(self.decoder, "density", dev),
                         runtime.get_sf3d_head_decoder_pack(self.decoder, "vertex_offset", dev)],
                        n_out=(1, 3), exp_act=(True, False), out_bias=(float(dens_spec.out_bias), 0.0),
                        axis_u=self._lattice_axis_u(dev), spatial_dim=h.lattice[1], align_corners=True,
                        out_sub=(float(self.cfg.isosurface_threshold), 0.0),  # density - threshold (:155) in the kernel's epilogue
                    )
                    level_done = True
                    if float(offs.out_bias) != 0.0:
                        deform = deform + float(offs.out_bias)
                elif tc:
                    dec = runtime.query_points_tc(
                        planes, runtime.get_sf3d_points_pack(self.decoder, dev), grid_vertices, self.cfg.radius,
                        float(dens_spec.out_bias), align_corners=True, sigmoid_vec=False, want=("out0_act", "vec"),
                    )
                    density, deform = dec["out0_act"], dec["vec"]
                else:
                    dec = runtime.sf3d_query(
                        planes, runtime.get_sf3d_heads(self.decoder, dev), float(dens_spec.out_bias), self.cfg.radius,
                        positions=grid_vertices, want=("density_act", "vertex_offset"),
                    )
                    density, deform = dec["density_act"], dec["vertex_offset"]
            else:
                values = self.query_triplane(grid_vertices, triplane)
                decoded = self.decoder(values, include=["vertex_offset", "density"])
                density, deform = decoded["density"], decoded["vertex_offset"].squeeze(0)
            sdf = density if level_done else density - self.cfg.isosurface_threshold  # :155
            # scale_tensor(mesh.v_pos, points_range, bbox) (:162-164) runs inside the vertex kernel: the same fp32 operations
            mesh = h(sdf.view(-1, 1), deform.view(-1, 3) if deform is not None else None, v_pos_affine=self._v_pos_affine(dev))
            meshes.append(mesh)
        return meshes
