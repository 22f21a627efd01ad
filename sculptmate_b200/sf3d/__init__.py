"""Host-side mirror of the reference's ``sf3d`` package for the mesh-extraction path
(/root/reference/StableFast/sf3d/{system.py, models/network.py, models/isosurface.py,
models/mesh.py}): same names, argument meaning and return types; the arithmetic runs in
the CUDA library (csrc/sf3d.cu)."""
from .models.isosurface import IsosurfaceHelper, MarchingTetrahedraHelper  # noqa: F401
from .models.mesh import Mesh  # noqa: F401
from .models.network import HeadSpec, MaterialMLP  # noqa: F401
from .system import SF3D  # noqa: F401
from .tets import kuhn_tet_grid, save_tet_grid  # noqa: F401
from .texture_baker import TextureBaker  # noqa: F401
