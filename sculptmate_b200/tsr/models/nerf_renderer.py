"""Drop-in for the query path of /root/reference/TripoSR/tsr/models/nerf_renderer.py.

``query_triplane(decoder, positions, triplane)`` keeps the reference signature and
return dict (nerf_renderer.py:41-91) but runs as one fused CUDA launch: the three
bilinear plane gathers, the concat, the whole NeRFMLP chain and the exp / sigmoid
tails, with no (N,120) feature tensor, no per-chunk launches and no torch.cat.
``query_lattice`` is the lattice specialisation extract_mesh uses (tensor cores).
``forward``/``_forward`` (:93-172, the volume renderer; SURVEY 8f rank 4 -- no caller in the add-on) keep the
reference's signature too: ray/box intersection with the reference's torch ops, sample positions and alpha
compositing in CUDA (``csrc/render.cu``), the field query in between on the tensor-core points kernel.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import ctypes
import weakref

import torch

from ... import _capi, runtime
from ..utils import BaseModule, rays_intersect_bbox


class TriplaneNeRFRenderer(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        radius: float

        feature_reduction: str = "concat"
        density_activation: str = "trunc_exp"
        density_bias: float = -1.0
        color_activation: str = "sigmoid"
        num_samples_per_ray: int = 128
        randomized: bool = False

    cfg: Config

    def configure(self) -> None:
        assert self.cfg.feature_reduction in ["concat", "mean"]
        self.chunk_size = 0
        # "fp32": CUDA-core kernel, reference precision (default); "tc": tcgen05 kernel, fp16 operands
        self.point_precision = "fp32"
        # prepared planes of the LAST triplane seen, valid only for that very tensor object (weak reference) at the
        # same _version: an address-based key would alias the next scene code the caching allocator hands the freed block
        self._scene_ref = None
        self._scene_key = None
        self._scene: Optional[runtime.ScenePlanes] = None

    def set_chunk_size(self, chunk_size: int):
        # kept for API compatibility (generate.py:25); the fused kernels do not chunk
        assert chunk_size >= 0, "chunk_size must be a non-negative integer (0 for no chunking)."
        self.chunk_size = chunk_size

    def _check_supported(self) -> None:
        c = self.cfg
        if c.feature_reduction != "concat" or c.density_activation != "exp" or c.color_activation != "sigmoid":
            raise NotImplementedError(
                "the CUDA path implements the TripoSR renderer config (config.yaml:32-37): "
                "feature_reduction=concat, density_activation=exp, color_activation=sigmoid; got "
                f"{c.feature_reduction}/{c.density_activation}/{c.color_activation}"
            )

    def _planes(self, decoder: torch.nn.Module, triplane: torch.Tensor):
        pack = runtime.get_decoder_pack(decoder, triplane.device)
        key = (triplane._version, triplane.data_ptr(), tuple(triplane.shape), pack.key)
        same = self._scene_ref is not None and self._scene_ref() is triplane and self._scene_key == key
        if not same:
            self._scene = runtime.prepare_scene(triplane, pack)
            self._scene_key = key
            self._scene_ref = weakref.ref(triplane)
        return pack, self._scene

    def query_triplane(
        self,
        decoder: torch.nn.Module,
        positions: torch.Tensor,
        triplane: torch.Tensor,
        precision: Optional[str] = None,
    ) -> Dict[str, torch.Tensor]:
        self._check_supported()
        input_shape = positions.shape[:-1]
        pack, scene = self._planes(decoder, triplane)
        precision = precision or self.point_precision
        if precision == "tc":
            tc = runtime.get_tsr_points_pack(decoder, triplane.device)
            r = runtime.query_points_tc(
                scene, tc, positions.reshape(-1, 3), self.cfg.radius, self.cfg.density_bias, align_corners=False, sigmoid_vec=True,
                want=("out0_raw", "out0_act", "vec", "vec_act"),
            )
            out = {"density": r["out0_raw"] - self.cfg.density_bias, "features": r["vec"], "density_act": r["out0_act"], "color": r["vec_act"]}
        elif precision == "fp32":
            out = runtime.query_points(scene, pack, positions.reshape(-1, 3), self.cfg.radius, self.cfg.density_bias)
        else:
            raise ValueError(f"precision must be 'fp32' or 'tc', got {precision!r}")
        return {k: v.view(*input_shape, v.shape[-1]) for k, v in out.items()}

    def query_lattice(
        self,
        decoder: torch.nn.Module,
        triplane: torch.Tensor,
        resolution: int,
        axis_u: Optional[torch.Tensor] = None,
        x_begin: int = 0,
        nx: Optional[int] = None,
        precision: str = "tc",
        out: Optional[torch.Tensor] = None,
        mc_signs=None,
    ) -> torch.Tensor:
        """density_act of query_triplane on the MarchingCubeHelper lattice, planes
        [x_begin, x_begin+nx) -> (nx,R,R); positions are generated in-kernel.
        ``mc_signs=(sub, sign)``: also leave the marching-cubes sign masks in the MC workspace
        (see ``runtime.query_lattice``)."""
        self._check_supported()
        pack, scene = self._planes(decoder, triplane)
        if axis_u is None:
            axis_u = runtime.lattice_axis(resolution, self.cfg.radius, device=triplane.device)
        return runtime.query_lattice(
            scene, pack, axis_u, resolution, self.cfg.radius, self.cfg.density_bias,
            x_begin=x_begin, nx=nx, precision=precision, out=out, mc_signs=mc_signs,
        )

    # ------------------------------------------------------------------ volume rendering
    def _forward(self, decoder: torch.nn.Module, triplane: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor,
                 precision: str = "tc", **kwargs) -> torch.Tensor:
        """comp_rgb (*rays_shape, 3) for one scene code (nerf_renderer.py:93-152).

        Like the reference this needs every ray to hit the box: there ``z_vals`` is built from the valid rays
        only and then added to ALL ray origins (:106-117), which raises a RuntimeError as soon as one ray
        misses; the same exception type is raised here."""
        self._check_supported()
        runtime._require_cuda(rays_o, "rays_o")
        rays_shape = rays_o.shape[:-1]
        dev = triplane.device
        o = rays_o.detach().to(torch.float32).reshape(-1, 3).contiguous()
        d = rays_d.detach().to(torch.float32).reshape(-1, 3).contiguous()
        n_rays = o.shape[0]
        t_near, t_far, valid = rays_intersect_bbox(o, d, self.cfg.radius)
        if not bool(valid.all()):
            raise RuntimeError(
                f"{int((~valid).sum())} of {n_rays} rays miss the +-{self.cfg.radius} box; TriplaneNeRFRenderer._forward "
                "(nerf_renderer.py:106-117) only supports rays that all intersect it"
            )
        S = int(self.cfg.num_samples_per_ray)
        t_vals = torch.linspace(0, 1, S + 1)  # :108 (values k/S; built on the host, moved once)
        t_mid = ((t_vals[:-1] + t_vals[1:]) / 2.0).to(dev)
        deltas = (t_vals[1:] - t_vals[:-1]).to(dev)  # :123 (the reference integrates over t in [0,1], not metric length)
        xyz = torch.empty((n_rays, S, 3), dtype=torch.float32, device=dev)
        comp = torch.empty((n_rays, 3), dtype=torch.float32, device=dev)
        lib = _capi.load()
        with torch.cuda.device(dev):
            st = runtime._stream_ptr(dev)
            _capi.check(lib.smb_ray_sample_positions(o.data_ptr(), d.data_ptr(), t_near.contiguous().data_ptr(), t_far.contiguous().data_ptr(),
                                                     t_mid.data_ptr(), n_rays, S, xyz.data_ptr(), st), "smb_ray_sample_positions")
            out = self.query_triplane(decoder, xyz, triplane, precision=precision)
            _capi.check(lib.smb_ray_composite(out["density_act"].contiguous().data_ptr(), out["color"].contiguous().data_ptr(), deltas.data_ptr(),
                                              n_rays, S, comp.data_ptr(), None, st), "smb_ray_composite")
        return comp.view(*rays_shape, 3)

    def forward(self, decoder: torch.nn.Module, triplane: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor,
                precision: str = "tc") -> torch.Tensor:
        if triplane.ndim == 4:
            return self._forward(decoder, triplane, rays_o, rays_d, precision=precision)
        return torch.stack([self._forward(decoder, triplane[i], rays_o[i], rays_d[i], precision=precision) for i in range(triplane.shape[0])], dim=0)

    def train(self, mode=True):
        self.randomized = mode and self.cfg.randomized
        return super().train(mode=mode)

    def eval(self):
        self.randomized = False
        return super().eval()
