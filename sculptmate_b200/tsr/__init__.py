"""Host-side mirror of the reference's ``tsr`` package for the extract_mesh path.

Same names, argument meaning and error behaviour as
/root/reference/TripoSR/tsr/{system.py, utils.py, models/*}; the arithmetic runs in
the CUDA library.
"""
from .models.isosurface import IsosurfaceHelper, MarchingCubeHelper  # noqa: F401
from .models.nerf_renderer import TriplaneNeRFRenderer  # noqa: F401
from .models.network_utils import NeRFMLP  # noqa: F401
from .system import TSR  # noqa: F401
from .utils import chunk_batch, get_activation, scale_tensor  # noqa: F401
