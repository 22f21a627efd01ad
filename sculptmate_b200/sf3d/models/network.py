"""Drop-in for MaterialMLP in /root/reference/StableFast/sf3d/models/network.py:139-208.

Parameter container with the reference's state-dict keys (``heads.<name>.{0,2,4}.*``).
``forward`` keeps the reference signature (``include`` / ``exclude``); the two heads the
mesh path uses (``density``: trunc_exp(head + out_bias), ``vertex_offset``) run in the CUDA
library, any other head (texture / material, out of scope) through its ``nn.Sequential``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import runtime
from ...tsr.utils import BaseModule


@dataclass
class HeadSpec:
    name: str
    out_channels: int
    n_hidden_layers: int
    output_activation: Optional[str] = None
    out_bias: float = 0.0


# eps of the reference's normalize() helper per dtype (sf3d/models/utils.py:57-76): F.normalize's own default
# (1e-12) is NOT what the shipped perturb_normal head uses
_NORMALIZE_EPS = {torch.float16: 1e-4, torch.bfloat16: 1e-4, torch.float32: 1e-7, torch.float64: 1e-8}


def _normalize(x: torch.Tensor, dim: int = -1) -> torch.Tensor:
    return F.normalize(x, dim=dim, p=2, eps=_NORMALIZE_EPS[x.dtype])


def _lin2srgb(x: torch.Tensor) -> torch.Tensor:
    return torch.where(x > 0.0031308, torch.pow(torch.clamp(x, min=0.0031308), 1.0 / 2.4) * 1.055 - 0.055, 12.92 * x).clamp(0.0, 1.0)


# name -> callable, the full table of network.py:98-136 (forward semantics; trunc_exp's forward is exp, its clamp only
# acts in backward, :85-92)
_ACTIVATIONS = {
    "none": lambda x: x, "linear": lambda x: x, "identity": lambda x: x,
    "lin2srgb": _lin2srgb,
    "exp": torch.exp, "trunc_exp": torch.exp,
    "shifted_exp": lambda x: torch.exp(x - 1.0), "shifted_trunc_exp": lambda x: torch.exp(x - 1.0),
    "sigmoid": torch.sigmoid, "tanh": torch.tanh,
    "shifted_softplus": lambda x: F.softplus(x - 1.0),
    "scale_-11_01": lambda x: x * 0.5 + 0.5,
    "negative": lambda x: -x,
    "normalize_channel_last": _normalize,
    "normalize_channel_first": lambda x: _normalize(x, dim=1),
}


def get_activation(name) -> Callable:
    """network.py:98-136: every name the reference knows, else torch.nn.functional.<name>, else ValueError."""
    if name is None:
        return lambda x: x
    name = name.lower()
    if name in _ACTIVATIONS:
        return _ACTIVATIONS[name]
    try:
        return getattr(F, name)
    except AttributeError:
        raise ValueError(f"Unknown activation function: {name}")


class MaterialMLP(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        in_channels: int = 120
        n_neurons: int = 64
        activation: str = "silu"
        heads: List[HeadSpec] = field(default_factory=lambda: [])

    cfg: Config

    def configure(self) -> None:
        self.cfg.heads = [h if isinstance(h, HeadSpec) else HeadSpec(**dict(h)) for h in self.cfg.heads]
        assert len(self.cfg.heads) > 0
        heads = {}
        for head in self.cfg.heads:
            layers = []
            for i in range(head.n_hidden_layers):
                layers += [
                    nn.Linear(self.cfg.in_channels if i == 0 else self.cfg.n_neurons, self.cfg.n_neurons),
                    self.make_activation(self.cfg.activation),
                ]
            layers += [nn.Linear(self.cfg.n_neurons, head.out_channels)]
            heads[head.name] = nn.Sequential(*layers)
        self.heads = nn.ModuleDict(heads)

    def make_activation(self, activation):
        if activation == "relu":
            return nn.ReLU(inplace=True)
        elif activation == "silu":
            return nn.SiLU(inplace=True)
        raise NotImplementedError

    def keys(self):
        return self.heads.keys()

    def head_spec(self, name: str) -> HeadSpec:
        for h in self.cfg.heads:
            if h.name == name:
                return h
        raise KeyError(name)

    def cuda_heads_supported(self) -> bool:
        """True when ``density`` and ``vertex_offset`` have the shipped architecture
        (config.yaml:45-65): 120 -> 64 -> 64 -> {1,3}, SiLU, density trunc_exp."""
        try:
            d, v = self.head_spec("density"), self.head_spec("vertex_offset")
        except KeyError:
            return False
        return (
            self.cfg.in_channels == 120 and self.cfg.n_neurons == 64 and self.cfg.activation == "silu"
            and d.n_hidden_layers == 2 and d.out_channels == 1 and (d.output_activation or "").lower() in ("trunc_exp", "exp")
            and v.n_hidden_layers == 2 and v.out_channels == 3 and (v.output_activation or "none").lower() in ("none", "linear", "identity")
            and float(v.out_bias) == 0.0
        )

    def forward(self, x, include: Optional[List] = None, exclude: Optional[List] = None):
        if include is not None and exclude is not None:
            raise ValueError("Cannot specify both include and exclude.")
        if include is not None:
            heads = [h for h in self.cfg.heads if h.name in include]
        elif exclude is not None:
            heads = [h for h in self.cfg.heads if h.name not in exclude]
        else:
            heads = self.cfg.heads
        out = {}
        fused = [h.name for h in heads if h.name in ("density", "vertex_offset")]
        if fused and x.is_cuda and self.cuda_heads_supported():
            lead = x.shape[:-1]
            res = runtime.sf3d_query(
                None, runtime.get_sf3d_heads(self, x.device), float(self.head_spec("density").out_bias), 1.0,
                features=x.reshape(-1, x.shape[-1]), want=tuple("density_act" if n == "density" else n for n in fused),
            )
            if "density" in fused:
                out["density"] = res["density_act"].view(*lead, 1)
            if "vertex_offset" in fused:
                out["vertex_offset"] = res["vertex_offset"].view(*lead, 3)
        for head in heads:
            if head.name not in out:  # heads outside the mesh path (texture / material): torch ops on the GPU
                if not x.is_cuda:
                    raise RuntimeError("sculptmate_b200: MaterialMLP needs CUDA tensors; the B200 path has no CPU fallback")
                out[head.name] = get_activation(head.output_activation)(self.heads[head.name](x) + head.out_bias)
        return {h.name: out[h.name] for h in heads}
