"""The mesh container ``triplane_to_meshes`` returns, mirroring the constructor of
/root/reference/StableFast/sf3d/models/mesh.py:19-37.  Normals, tangents, remeshing and
UV unwrapping (mesh.py:38-277) are post-mesh CPU work and stay in the reference."""
from __future__ import annotations

from typing import Any, Dict

import torch


class Mesh:
    def __init__(self, v_pos: torch.Tensor, t_pos_idx: torch.Tensor, **kwargs) -> None:
        self.v_pos = v_pos
        self.t_pos_idx = t_pos_idx
        self.extras: Dict[str, Any] = {}
        for k, v in kwargs.items():
            self.add_extra(k, v)

    def add_extra(self, k, v) -> None:
        self.extras[k] = v

    @property
    def requires_grad(self):
        return self.v_pos.requires_grad
