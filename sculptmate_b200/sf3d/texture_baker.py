"""Drop-in for /root/reference/StableFast/sf3d/texture_baker/baker.py (``TextureBaker``): UV-space rasterisation and
barycentric interpolation for the texture bake of ``SF3D.generate_mesh`` (sf3d/system.py:361-389).

The reference marshals everything to numpy and calls ``rasterize_cpu`` / ``interpolate_cpu`` in a Windows-only DLL
(baker.py:30-57,92-118); here both run as CUDA kernels (``csrc/bake.cu``) on tensors that stay on the GPU, restating the
Python functions of the same name that ship beside the DLL (texture_baker/common.py).  Same method names, argument
order and return shapes; there is no CPU path."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from .. import _capi
from ..runtime import _stream_ptr


def _dev(device, *tensors) -> torch.device:
    d = torch.device(device) if device is not None else tensors[0].device
    if d.type != "cuda":
        raise RuntimeError("sculptmate_b200: TextureBaker needs a CUDA device; the B200 path has no CPU fallback")
    return d


class TextureBaker(nn.Module):
    def __init__(self):
        super().__init__()
        self._ws = {}

    def rasterize(self, uv: Tensor, face_indices: Tensor, bake_resolution: int, device=None) -> Tensor:
        """(bake_resolution, bake_resolution, 4) fp32: barycentrics + triangle index per texel, -1 where empty (baker.py:12-57)."""
        dev = _dev(device, uv)
        uv = uv.detach().to(device=dev, dtype=torch.float32).contiguous()
        faces = face_indices.detach().to(device=dev, dtype=torch.int32).contiguous()  # baker.py:44
        res = int(bake_resolution)
        rast = torch.empty((res, res, 4), dtype=torch.float32, device=dev)
        lib = _capi.load()
        key = (str(dev), res)
        if key not in self._ws:
            self._ws = {key: torch.empty(lib.smb_bake_workspace_bytes(res), dtype=torch.uint8, device=dev)}
        ws = self._ws[key]
        with torch.cuda.device(dev):
            _capi.check(lib.smb_bake_rasterize(uv.data_ptr(), faces.data_ptr(), int(uv.shape[0]), int(faces.shape[0]), res, rast.data_ptr(),
                                               ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "smb_bake_rasterize")
        return rast

    def get_mask(self, rast: Tensor) -> Tensor:
        return rast[..., -1] >= 0  # baker.py:59-69

    def interpolate(self, attr: Tensor, rast: Tensor, face_indices: Tensor, bake_resolution: int, device=None) -> Tensor:
        """(bake_resolution, bake_resolution, C) fp32: attributes interpolated with the rasterised barycentrics (baker.py:71-118)."""
        dev = _dev(device, rast)
        attr = attr.detach().to(device=dev, dtype=torch.float32).contiguous()
        faces = face_indices.detach().to(device=dev, dtype=torch.int32).contiguous()
        rast = rast.detach().to(device=dev, dtype=torch.float32).contiguous()
        res, C = int(bake_resolution), int(attr.shape[1])
        if rast.shape != (res, res, 4):
            raise ValueError(f"rast must be ({res},{res},4), got {tuple(rast.shape)}")
        out = torch.empty((res, res, C), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _capi.check(_capi.load().smb_bake_interpolate(attr.data_ptr(), C, faces.data_ptr(), int(faces.shape[0]), rast.data_ptr(), res,
                                                          out.data_ptr(), _stream_ptr(dev)), "smb_bake_interpolate")
        return out

    def forward(self, attr: Tensor, uv: Tensor, face_indices: Tensor, bake_resolution: int, device=None) -> Tensor:
        """rasterize + interpolate.  (The reference's forward passes a stray ``uv`` to interpolate, baker.py:140-141, and
        raises TypeError; nothing calls it.  This one returns the baked texture its docstring promises.)"""
        rast = self.rasterize(uv, face_indices, bake_resolution, device)
        return self.interpolate(attr, rast, face_indices, bake_resolution, device)
