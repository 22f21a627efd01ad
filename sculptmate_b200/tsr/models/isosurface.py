"""Drop-in for /root/reference/TripoSR/tsr/models/isosurface.py (MarchingCubeHelper).

``forward`` keeps the reference contract -- (R^3) or (R^3,1) level in, (V,3) float
vertices in [0,1] and (F,3) long faces out, on ``level.device`` -- but the surface is
extracted by the CUDA pipeline (csrc/mcubes.cu) instead of a D2H copy + CPU
``skimage.measure.marching_cubes``.  The algorithm is a classic 256-case marching
cubes with a canonical deterministic ordering, not Lewiner's (MC
parity against scikit-image is unpinned).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from ... import runtime
from ..._capi import MC_DIV, MC_FLIP


class IsosurfaceHelper(nn.Module):
    points_range: Tuple[float, float] = (0, 1)

    @property
    def grid_vertices(self) -> torch.FloatTensor:
        raise NotImplementedError


class MarchingCubeHelper(IsosurfaceHelper):
    def __init__(self, resolution: int) -> None:
        super().__init__()
        self.resolution = resolution
        self._grid_vertices: Optional[torch.FloatTensor] = None

    @property
    def grid_vertices(self) -> torch.FloatTensor:
        """(R^3,3) lattice, x slowest / z fastest (isosurface.py:25-39).  Kept for API
        compatibility; the fused extract_mesh path never materialises it."""
        if self._grid_vertices is None:
            ax = torch.linspace(*self.points_range, self.resolution)
            x, y, z = torch.meshgrid(ax, ax, ax, indexing="ij")
            self._grid_vertices = torch.stack([x.reshape(-1), y.reshape(-1), z.reshape(-1)], dim=-1)
        return self._grid_vertices

    def forward(self, level: torch.FloatTensor) -> Tuple[torch.FloatTensor, torch.LongTensor]:
        R = self.resolution
        # isosurface.py:45 negates the input and extracts the 0-level set; the kernel
        # applies val = (grid - 0) * (-1) on the fly instead of writing a negated copy.
        grid = level.detach().to(torch.float32).contiguous().view(R, R, R)
        # isosurface.py:52-53: faces[:, [1,0,2]] and verts / (R - 1)
        v_pos, t_pos_idx, pend = runtime.mc_extract(grid, sub=0.0, sign=-1.0, flags=MC_FLIP | MC_DIV, vdiv=float(R - 1.0))
        if pend.nverts == 0 or pend.ntris == 0:
            runtime.raise_for_empty_surface(grid, 0.0, -1.0)
        return v_pos, t_pos_idx
