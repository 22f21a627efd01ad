"""Mesh hand-off: sculptmate_b200.tsr.blender_io against the reference's own
TSR.import_obj_blender (tsr/system.py:127-168), both driven through a stand-in ``bpy`` (tests/fake_bpy.py)."""
import sys

import numpy as np
import pytest
import torch

import fake_bpy


def _mesh(seed=0, V=40, F=90):
    rng = np.random.RandomState(seed)
    verts = rng.rand(V, 3).astype(np.float32)
    faces = rng.randint(0, V, size=(F, 3)).astype(np.int64)
    colors = rng.rand(V, 3).astype(np.float32)
    return verts, faces, colors


def _run_ours(verts, faces, colors, monkeypatch, **kw):
    from sculptmate_b200.tsr import blender_io

    bpy = fake_bpy.make()
    monkeypatch.setitem(sys.modules, "bpy", bpy)
    blender_io.import_obj_blender(verts, faces, colors, name="m", **kw)
    return fake_bpy.snapshot(bpy)


def _same(a, b):
    (ma, la, oa), (mb, lb, ob) = a, b
    assert la == lb and oa == ob and len(ma) == len(mb)
    for x, y in zip(ma, mb):
        assert x["name"] == y["name"] and x["polys"] == y["polys"] and x["materials"] == y["materials"]
        np.testing.assert_array_equal(x["verts"], y["verts"])
        np.testing.assert_array_equal(x["loops"], y["loops"])
        assert x["layers"].keys() == y["layers"].keys()
        for k in x["layers"]:
            np.testing.assert_array_equal(x["layers"][k], y["layers"][k])


@pytest.mark.parametrize("with_colors", [True, False])
def test_sink_equals_reference_import_obj_blender(monkeypatch, with_colors):
    from oracle import ref_shim

    if not ref_shim.reference_available():
        pytest.skip("reference checkout not present")
    ref_shim.load_triposr()
    import tsr.system as ref_system  # the reference's own module

    verts, faces, colors = _mesh()
    colors = colors if with_colors else None
    bpy = fake_bpy.make()
    monkeypatch.setattr(ref_system, "bpy", bpy)
    ref_system.TSR.import_obj_blender(None, verts, faces, None if colors is None else colors.copy(), name="m")
    ref = fake_bpy.snapshot(bpy)
    ours = _run_ours(verts, faces, colors, monkeypatch)
    _same(ref, ours)
    if with_colors:
        layer = ours[0][0]["layers"]["m_VC"]
        assert layer.shape == (3 * len(faces), 4) and np.all(layer[:, 3] == 1.0)
        # a pre-gathered loop_colors array (what TSR.extract_mesh passes) gives the same layer
        lc = np.hstack((colors, np.ones((len(colors), 1), np.float32)))[faces.reshape(-1)]
        _same(ref, _run_ours(verts, faces, colors, monkeypatch, loop_colors=lc))


def test_tsr_passes_loop_colors_only_to_sinks_that_take_them():
    from sculptmate_b200.tsr import TSR

    m = TSR()
    assert not m._sink_takes_loop_colors()
    m.mesh_sink = lambda verts, faces, vertex_colors=None, name="x": None
    assert not m._sink_takes_loop_colors()
    from sculptmate_b200.tsr import blender_io

    m.mesh_sink = blender_io.import_obj_blender
    assert m._sink_takes_loop_colors()


@pytest.mark.gpu
@pytest.mark.parametrize("V,F", [(40, 90), (1, 1), (100003, 250007), (5, 0)])
def test_loop_colors_and_int32_faces_on_gpu(V, F):
    from sculptmate_b200.tsr import blender_io

    verts, faces, colors = _mesh(1, V, F)
    fc, cc = torch.from_numpy(faces).cuda(), torch.from_numpy(colors).cuda()
    lc = blender_io.loop_colors(cc, fc).cpu().numpy()
    ref = np.hstack((colors, np.ones((V, 1), np.float32)))[faces.reshape(-1)].reshape(-1, 4)
    np.testing.assert_array_equal(lc, ref)
    f32 = blender_io.faces_int32(fc)
    assert f32.dtype == torch.int32
    np.testing.assert_array_equal(f32.cpu().numpy(), faces.astype(np.int32))
    if F:
        bad = fc.clone()
        bad[0, 0] = V
        with pytest.raises(IndexError):
            blender_io.loop_colors(cc, bad)


@pytest.mark.gpu
def test_extract_mesh_hands_loop_colors_to_the_blender_sink(golden, monkeypatch):
    from sculptmate_b200.tsr import blender_io
    from test_gpu_field import _model

    g = golden("extract_mesh.npz")
    m = _model(g)
    bpy = fake_bpy.make()
    monkeypatch.setitem(sys.modules, "bpy", bpy)
    m.mesh_sink = blender_io.import_obj_blender
    tp = torch.from_numpy(g["triplane"]).cuda()
    m.extract_mesh(tp[None], enable_texture=True, mesh_name="g", resolution=24, threshold=float(g["threshold"]))
    meshes, linked, _ = fake_bpy.snapshot(bpy)
    assert linked == ["g"] and len(meshes) == 1
    layer = meshes[0]["layers"]["g_VC"]
    # the layer equals what the reference's per-loop loop would assign from the (V,3) colours
    col = m.renderer.query_triplane(m.decoder, torch.from_numpy(meshes[0]["verts"].astype(np.float32)).cuda(), tp, precision="tc")["color"].cpu().numpy()
    np.testing.assert_allclose(layer[:, :3], col[meshes[0]["loops"]], atol=1e-6)
    assert np.all(layer[:, 3] == 1.0)
