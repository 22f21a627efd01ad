"""Drop-in for NeRFMLP in /root/reference/TripoSR/tsr/models/network_utils.py:35-124.

Parameter container with the reference's state-dict keys
(``layers.{0,2,...}.{weight,bias}``) so checkpoints load unchanged.  On the hot path
the renderer never calls ``forward``: it hands these parameters to the fused CUDA
kernels (``runtime.get_decoder_pack``).  ``forward`` on pre-computed features runs the
decoder in the CUDA library too (fp32 kernel); there is no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from ..utils import BaseModule


class NeRFMLP(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        in_channels: int
        n_neurons: int
        n_hidden_layers: int
        activation: str = "relu"
        bias: bool = True
        weight_init: Optional[str] = "kaiming_uniform"
        bias_init: Optional[str] = None

    cfg: Config

    def configure(self) -> None:
        c = self.cfg
        dims = [c.in_channels] + [c.n_neurons] * c.n_hidden_layers
        layers = []
        for d_in, d_out in zip(dims[:-1], dims[1:]):
            layers += [self.make_linear(d_in, d_out, c.bias, c.weight_init, c.bias_init), self.make_activation(c.activation)]
        layers.append(self.make_linear(c.n_neurons, 4, c.bias, c.weight_init, c.bias_init))  # density 1 + features 3
        self.layers = nn.Sequential(*layers)

    def make_linear(self, dim_in, dim_out, bias=True, weight_init=None, bias_init=None):
        layer = nn.Linear(dim_in, dim_out, bias=bias)
        if weight_init == "kaiming_uniform":
            torch.nn.init.kaiming_uniform_(layer.weight, nonlinearity="relu")
        elif weight_init is not None:
            raise NotImplementedError
        if bias:
            if bias_init == "zero":
                torch.nn.init.zeros_(layer.bias)
            elif bias_init is not None:
                raise NotImplementedError
        return layer

    def make_activation(self, activation):
        if activation == "relu":
            return nn.ReLU(inplace=True)
        if activation == "silu":
            return nn.SiLU(inplace=True)
        raise NotImplementedError

    def forward(self, x):
        from ... import runtime

        lead = x.shape[:-1]
        out = runtime.decoder_forward(runtime.get_decoder_pack(self, x.device), x.reshape(-1, x.shape[-1]))
        return {"density": out["density"].view(*lead, 1), "features": out["features"].view(*lead, 3)}
