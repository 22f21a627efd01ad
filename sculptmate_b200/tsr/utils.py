"""Helpers on the extract_mesh path, mirroring /root/reference/TripoSR/tsr/utils.py.

Only the three helpers the path uses are here (T2, T4, T6) plus
a dependency-free ``BaseModule`` (the reference's needs omegaconf, utils.py:44-59).
"""
from __future__ import annotations

import dataclasses
from collections import defaultdict
from dataclasses import dataclass
from typing import Any, Callable, Dict, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

ValidScale = Union[Tuple[float, float], torch.FloatTensor]


def rays_intersect_bbox(rays_o: torch.Tensor, rays_d: torch.Tensor, radius, near: float = 0.0, valid_thresh: float = 0.01):
    """Slab test of rays against the (slightly tightened) +-radius box: the reference's utils.py:115-149 with the
    same torch ops in the same order (so t_near / t_far are bit-identical on the same device):
    returns ``t_near``, ``t_far`` (``*shape, 1``) and ``rays_valid`` (``*shape``)."""
    shape = rays_o.shape[:-1]
    o, d = rays_o.view(-1, 3), rays_d.view(-1, 3)
    d_safe = torch.where(d.abs() < 1e-6, torch.full_like(d, 1e-6), d)  # never divide by ~0
    if type(radius) in [int, float]:
        radius = torch.FloatTensor([[-radius, radius]] * 3).to(o.device)
    radius = (1.0 - 1.0e-3) * radius  # the hit points must lie inside the box
    hit_hi = (radius[..., 1] - o) / d_safe
    hit_lo = (radius[..., 0] - o) / d_safe
    t_near = torch.minimum(hit_hi, hit_lo).amax(dim=-1).clamp_min(near)
    t_far = torch.maximum(hit_hi, hit_lo).amin(dim=-1)
    valid = t_far - t_near > valid_thresh
    t_near[torch.where(~valid)] = 0.0
    t_far[torch.where(~valid)] = 0.0
    return t_near.view(*shape, 1), t_far.view(*shape, 1), valid.view(*shape)


def scale_tensor(dat: torch.FloatTensor, inp_scale: ValidScale, tgt_scale: ValidScale):
    """Affine remap, same operation order as utils.py:222-231."""
    if inp_scale is None:
        inp_scale = (0, 1)
    if tgt_scale is None:
        tgt_scale = (0, 1)
    if isinstance(tgt_scale, torch.FloatTensor):
        assert dat.shape[-1] == tgt_scale.shape[-1]
    dat = (dat - inp_scale[0]) / (inp_scale[1] - inp_scale[0])
    dat = dat * (tgt_scale[1] - tgt_scale[0]) + tgt_scale[0]
    return dat


def get_activation(name) -> Callable:
    """utils.py:234-252."""
    if name is None:
        return lambda x: x
    name = name.lower()
    if name == "none":
        return lambda x: x
    elif name == "exp":
        return lambda x: torch.exp(x)
    elif name == "sigmoid":
        return lambda x: torch.sigmoid(x)
    elif name == "tanh":
        return lambda x: torch.tanh(x)
    elif name == "softplus":
        return lambda x: F.softplus(x)
    else:
        try:
            return getattr(F, name)
        except AttributeError:
            raise ValueError(f"Unknown activation function: {name}")


def chunk_batch(func: Callable, chunk_size: int, *args, **kwargs) -> Any:
    """utils.py:152-216: slice the tensor arguments along dim 0, call, concatenate.

    The fused kernels do not need chunking (set_chunk_size stays a harmless knob); this
    is kept because callers outside the fused path use it with arbitrary functions.
    """
    if chunk_size <= 0:
        return func(*args, **kwargs)
    B = None
    for arg in list(args) + list(kwargs.values()):
        if isinstance(arg, torch.Tensor):
            B = arg.shape[0]
            break
    assert B is not None, "No tensor found in args or kwargs, cannot determine batch size."
    pieces = defaultdict(list)
    out_type = None
    n_items = 0
    for start in range(0, max(1, B), chunk_size):  # max(1, B): B == 0 still calls once
        sl = slice(start, start + chunk_size)
        res = func(
            *[a[sl] if isinstance(a, torch.Tensor) else a for a in args],
            **{k: a[sl] if isinstance(a, torch.Tensor) else a for k, a in kwargs.items()},
        )
        if res is None:
            continue
        out_type = type(res)
        if isinstance(res, torch.Tensor):
            res = {0: res}
        elif isinstance(res, (tuple, list)):
            n_items = len(res)
            res = dict(enumerate(res))
        elif not isinstance(res, dict):
            raise TypeError(f"Return value of func must be in type [torch.Tensor, list, tuple, dict], get {type(res)}.")
        for k, v in res.items():
            pieces[k].append(v if torch.is_grad_enabled() or v is None else v.detach())
    if out_type is None:
        return None
    merged: Dict[Any, Optional[torch.Tensor]] = {}
    for k, v in pieces.items():
        if all(x is None for x in v):
            merged[k] = None
        elif all(isinstance(x, torch.Tensor) for x in v):
            merged[k] = torch.cat(v, dim=0)
        else:
            raise TypeError(
                f"Unsupported types in return value of func: {[type(x) for x in v if not isinstance(x, torch.Tensor)]}"
            )
    if out_type is torch.Tensor:
        return merged[0]
    if out_type in (tuple, list):
        return out_type([merged[i] for i in range(n_items)])
    return merged


class _Cfg(dict):
    """dict with attribute access (stands in for omegaconf.DictConfig)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def parse_structured(fields: Any, cfg: Optional[dict] = None) -> _Cfg:
    """Dataclass defaults overlaid with ``cfg`` (utils.py:16-18 without omegaconf)."""
    out = _Cfg()
    required = []
    if dataclasses.is_dataclass(fields):
        for f in dataclasses.fields(fields):
            if f.default is not dataclasses.MISSING:
                out[f.name] = f.default
            elif f.default_factory is not dataclasses.MISSING:  # type: ignore[misc]
                out[f.name] = f.default_factory()  # type: ignore[misc]
            else:
                required.append(f.name)
    for k, v in dict(cfg or {}).items():
        out[k] = v
    missing = [k for k in required if k not in out]
    if missing:
        raise ValueError(f"missing mandatory config value(s): {missing}")
    return out


class BaseModule(nn.Module):
    @dataclass
    class Config:
        pass

    cfg: Config

    def __init__(self, cfg: Optional[dict] = None, *args, **kwargs) -> None:
        super().__init__()
        self.cfg = parse_structured(self.Config, cfg)
        self.configure(*args, **kwargs)

    def configure(self, *args, **kwargs) -> None:
        raise NotImplementedError
