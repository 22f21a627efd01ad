"""Mesh hand-off to Blender: a sink with the signature of the reference's
``TSR.import_obj_blender`` (/root/reference/TripoSR/tsr/system.py:127-168) that makes the same
``bpy`` calls, except that the per-loop Python assignment of vertex colours (:143-146, one
attribute write per polygon corner: 5.8 M of them for a 256^3 mesh) is one
``foreach_set("color", ...)`` over an array gathered on the GPU (``smb_mesh_loop_colors``).

``bpy`` is imported lazily: this module is importable outside Blender (the tests drive it with a
stand-in ``bpy``), and nothing here runs unless a caller installs the sink:

    fast.mesh_sink = sculptmate_b200.tsr.blender_io.import_obj_blender
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from .. import _capi
from ..runtime import _require_cuda, _stream_ptr


def loop_colors(vertex_colors: torch.Tensor, faces: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """(V,3) fp32 colours + (F,3) int64 faces, both on the GPU -> (3F,4) fp32 RGBA per polygon loop, in the
    order ``Mesh.from_pydata`` numbers the loops (3*f + corner).  Raises IndexError on an out-of-range face index
    (numpy fancy indexing, which this replaces, would too)."""
    _require_cuda(vertex_colors, "vertex_colors")
    _require_cuda(faces, "faces")
    if vertex_colors.dim() != 2 or vertex_colors.shape[1] != 3 or vertex_colors.dtype != torch.float32:
        raise ValueError("vertex_colors must be (V,3) float32")
    if faces.dim() != 2 or faces.shape[1] != 3 or faces.dtype != torch.int64:
        raise ValueError("faces must be (F,3) int64")
    vertex_colors, faces = vertex_colors.contiguous(), faces.contiguous()
    dev = faces.device
    out = torch.empty((3 * faces.shape[0], 4), dtype=torch.float32, device=dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _capi.check(
            _capi.load().smb_mesh_loop_colors(vertex_colors.data_ptr(), faces.data_ptr(), int(vertex_colors.shape[0]), int(faces.shape[0]),
                                              float(alpha), out.data_ptr(), bad.data_ptr(), _stream_ptr(dev)),
            "smb_mesh_loop_colors",
        )
    if int(bad.item()):
        raise IndexError("face index outside the vertex table")
    return out


def faces_int32(faces: torch.Tensor) -> torch.Tensor:
    """(F,3) int64 on the GPU -> (F,3) int32 (Blender's MeshLoop.vertex_index width)."""
    _require_cuda(faces, "faces")
    if faces.dim() != 2 or faces.shape[1] != 3 or faces.dtype != torch.int64:
        raise ValueError("faces must be (F,3) int64")
    faces = faces.contiguous()
    out = torch.empty(faces.shape, dtype=torch.int32, device=faces.device)
    with torch.cuda.device(faces.device):
        _capi.check(_capi.load().smb_mesh_faces_i32(faces.data_ptr(), int(faces.shape[0]), out.data_ptr(), _stream_ptr(faces.device)),
                    "smb_mesh_faces_i32")
    return out


def import_obj_blender(verts, faces, vertex_colors=None, name: str = "NewMesh", loop_colors: Optional[np.ndarray] = None):
    """Drop-in for ``TSR.import_obj_blender(self, verts, faces, vertex_colors, name)`` (system.py:127-168).

    Same objects, layer, material and node graph as the reference.  ``loop_colors`` ((3F,4) float32, from
    :func:`loop_colors` -- ``TSR.extract_mesh`` passes it when the sink accepts the keyword) replaces the
    reference's per-loop Python loop by one ``foreach_set``; without it the array is gathered with numpy
    from ``vertex_colors`` exactly as the reference indexes it."""
    import bpy  # noqa: PLC0415  (only available inside Blender)

    mesh_data = bpy.data.meshes.new(name=name)
    mesh_data.from_pydata(verts, [], faces)                                   # system.py:129
    new_object = bpy.data.objects.new(name=name, object_data=mesh_data)       # :130
    bpy.context.collection.objects.link(new_object)                           # :131
    if vertex_colors is None:
        return new_object
    if loop_colors is None:
        vc = np.asarray(vertex_colors)
        if vc.shape[1] == 3:                                                  # :133-135
            vc = np.hstack((vc, np.ones((vc.shape[0], 1))))
        loop_colors = vc[np.asarray(faces).reshape(-1)]                       # what :143-146 assigns, loop by loop
    vertex_colors_name = f"{name}_VC"                                         # :137
    mesh_data.vertex_colors.new(name=vertex_colors_name)
    color_layer = mesh_data.vertex_colors[vertex_colors_name]
    color_layer.data.foreach_set("color", np.ascontiguousarray(loop_colors, dtype=np.float32).reshape(-1))

    mat = bpy.data.materials.new(name="VertexColorMaterial")                  # :148-168, unchanged
    mesh_data.materials.append(mat)
    mat.use_nodes = True
    nodes = mat.node_tree.nodes
    links = mat.node_tree.links
    for node in list(nodes):
        nodes.remove(node)
    output_node = nodes.new(type="ShaderNodeOutputMaterial")
    principled_node = nodes.new(type="ShaderNodeBsdfPrincipled")
    vertex_color_node = nodes.new(type="ShaderNodeVertexColor")
    vertex_color_node.layer_name = vertex_colors_name
    links.new(vertex_color_node.outputs["Color"], principled_node.inputs["Base Color"])
    links.new(principled_node.outputs["BSDF"], output_node.inputs["Surface"])
    principled_node.inputs["Roughness"].default_value = 1
    principled_node.inputs["IOR"].default_value = 1.00
    return new_object
