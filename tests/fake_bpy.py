"""A minimal stand-in for Blender's ``bpy`` (TEST INFRASTRUCTURE): just the calls
TSR.import_obj_blender makes (/root/reference/TripoSR/tsr/system.py:127-168), with the data model those calls
rely on -- ``from_pydata`` numbers polygon loops consecutively (3*f + corner for triangles),
``vertex_colors[...].data[idx].color`` is per loop, ``foreach_set`` is the bulk form of the same assignment."""
from __future__ import annotations

import types

import numpy as np


class _Loop:
    def __init__(self, v):
        self.vertex_index = int(v)


class _Poly:
    def __init__(self, start, n):
        self.loop_indices = range(start, start + n)


class _ColorElem:
    def __init__(self):
        self.color = (0.0, 0.0, 0.0, 0.0)


class _ColorData(list):
    def foreach_set(self, attr, flat):
        assert attr == "color"
        flat = np.asarray(flat)
        assert flat.ndim == 1 and flat.size == 4 * len(self)
        for i, e in enumerate(self):
            e.color = tuple(flat[4 * i : 4 * i + 4])


class _ColorLayer:
    def __init__(self, name, nloops):
        self.name = name
        self.data = _ColorData(_ColorElem() for _ in range(nloops))


class _ColorLayers(dict):
    def __init__(self, mesh):
        super().__init__()
        self._mesh = mesh

    def new(self, name):
        self[name] = _ColorLayer(name, len(self._mesh.loops))
        return self[name]


class _Mesh:
    def __init__(self, name):
        self.name = name
        self.vertices, self.loops, self.polygons, self.materials = [], [], [], []
        self.vertex_colors = _ColorLayers(self)

    def from_pydata(self, verts, edges, faces):
        assert len(edges) == 0
        self.vertices = [tuple(float(x) for x in v) for v in verts]
        for f in faces:
            self.polygons.append(_Poly(len(self.loops), len(f)))
            self.loops.extend(_Loop(v) for v in f)


class _Socket:
    def __init__(self, node, name):
        self.node, self.name, self.default_value = node, name, None


class _Node:
    def __init__(self, type):  # noqa: A002
        self.type, self.layer_name = type, None
        self.inputs, self.outputs = _Sockets(self), _Sockets(self)


class _Sockets(dict):
    def __init__(self, node):
        super().__init__()
        self._node = node

    def __missing__(self, key):
        self[key] = _Socket(self._node, key)
        return self[key]


class _Nodes:
    def __init__(self):
        self._l = [_Node("ShaderNodeBsdfPrincipled"), _Node("ShaderNodeOutputMaterial")]  # what use_nodes creates

    def __iter__(self):
        return iter(list(self._l))  # Blender's collection tolerates removal while iterating (system.py:156-157)

    def __len__(self):
        return len(self._l)

    def remove(self, n):
        self._l.remove(n)

    def new(self, type):  # noqa: A002
        n = _Node(type)
        self._l.append(n)
        return n


class _Links(list):
    def new(self, a, b):
        self.append((a.node.type, a.name, b.node.type, b.name))


class _Material:
    def __init__(self, name):
        self.name, self.use_nodes = name, False
        self.node_tree = types.SimpleNamespace(nodes=_Nodes(), links=_Links())


class _Named(list):
    def __init__(self, factory):
        super().__init__()
        self._factory = factory

    def new(self, name, **kw):
        o = self._factory(name, **kw)
        self.append(o)
        return o


def make():
    """A fresh ``bpy`` module object."""
    bpy = types.ModuleType("bpy")
    linked = []
    bpy.data = types.SimpleNamespace(
        meshes=_Named(_Mesh),
        objects=_Named(lambda name, object_data=None: types.SimpleNamespace(name=name, data=object_data)),
        materials=_Named(_Material),
    )
    bpy.context = types.SimpleNamespace(collection=types.SimpleNamespace(objects=types.SimpleNamespace(link=linked.append)))
    bpy.linked = linked
    return bpy


def snapshot(bpy):
    """Everything import_obj_blender leaves behind, as plain data for comparison."""
    out = []
    for m in bpy.data.meshes:
        layers = {k: np.array([e.color for e in v.data], dtype=np.float64) for k, v in m.vertex_colors.items()}
        mats = []
        for mat in m.materials:
            nodes = [(n.type, n.layer_name, {k: s.default_value for k, s in n.inputs.items() if s.default_value is not None}) for n in mat.node_tree.nodes]
            mats.append((mat.name, mat.use_nodes, nodes, list(mat.node_tree.links)))
        out.append(dict(name=m.name, verts=np.array(m.vertices), loops=np.array([l.vertex_index for l in m.loops]),
                        polys=[(p.loop_indices.start, len(p.loop_indices)) for p in m.polygons], layers=layers, materials=mats))
    return out, [o.name for o in bpy.linked], [o.name for o in bpy.data.objects]
