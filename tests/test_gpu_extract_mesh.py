"""GPU parity of the whole path behind the reference's extract_mesh signature."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import golden_decoder, mesh_topology
from test_gpu_field import _model

pytestmark = pytest.mark.gpu
RADIUS = 0.87


def test_extract_mesh_dropin_against_reference_golden(golden):
    g = golden("extract_mesh.npz")
    m = _model(g)
    R, thr = int(g["resolution"]), float(g["threshold"])
    tp = torch.from_numpy(g["triplane"]).cuda()
    ret = m.extract_mesh(tp[None], enable_texture=True, mesh_name="golden", resolution=R, threshold=thr)
    assert ret is None  # the reference returns None and delivers through import_obj_blender
    verts, faces, colors, name = m.meshes[-1]
    assert name == "golden" and verts.dtype == np.float32 and faces.dtype == np.int64 and colors.dtype == np.float32
    assert verts.shape[1] == 3 and faces.shape[1] == 3 and colors.shape == verts.shape
    assert np.abs(verts).max() <= RADIUS + 1e-6
    # tensor-core density differs from the reference's fp32 density in the 4th digit, so a few
    # cells may flip: counts within 1 %, geometry close
    assert abs(len(verts) / len(g["verts"]) - 1) < 0.01 and abs(len(faces) / len(g["faces"]) - 1) < 0.01
    # with the fp32 kernel the cube cases are identical -> connectivity identical to the golden
    v32, f32 = m.extract_mesh_tensors(tp, R, thr, precision="fp32")
    np.testing.assert_array_equal(f32.cpu().numpy(), g["faces"])
    assert np.abs(v32.cpu().numpy() - g["verts"]).max() < 1e-4
    # the reference-shaped slow path (grid_vertices -> query_triplane -> helper) agrees too
    vu, fu = m.extract_mesh_unfused(tp, R, thr)
    np.testing.assert_array_equal(fu.cpu().numpy(), g["faces"])
    assert np.abs(vu.cpu().numpy() - g["verts"]).max() < 1e-4


def test_colour_query_at_golden_vertices(golden):
    g = golden("extract_mesh.npz")
    m = _model(g)
    tp = torch.from_numpy(g["triplane"]).cuda()
    col = m.renderer.query_triplane(m.decoder, torch.from_numpy(g["verts"]).cuda(), tp)["color"]
    assert np.abs(col.cpu().numpy() - g["colors"]).max() < 2e-5


@pytest.mark.parametrize("R", [32, 64, 128])
def test_mesh_bit_exact_vs_oracle_on_shared_density(golden, R):
    """north_star: connectivity bit-exact when both sides are fed the same density grid."""
    from oracle import mc_oracle

    g = golden("extract_mesh.npz")
    m = _model(g)
    tp = torch.from_numpy(g["triplane"]).cuda()
    dens = m.renderer.query_lattice(m.decoder, tp, R)
    thr = float(dens.median())
    v, f = m.extract_mesh_tensors(tp, R, thr)
    v_ref, f_ref, _ = mc_oracle.marching_cubes_slab(
        dens.cpu().numpy(), sub=np.float32(thr), flags=7, vdiv=float(R - 1.0), vmul=float(RADIUS - (-RADIUS)), vadd=float(-RADIUS)
    )
    np.testing.assert_array_equal(v.cpu().numpy(), v_ref)
    np.testing.assert_array_equal(f.cpu().numpy(), f_ref)


@pytest.mark.parametrize("R", [17, 33, 100, 130, 160])
def test_fused_sign_masks_equal_the_classification_pass(golden, R):
    """The lattice kernel ballots the marching-cubes case bits while the densities are in registers
    (smb_query_lattice_tc_signs); the masks and the mesh must equal what the stand-alone pass over the
    stored grid gives, also when R is not a multiple of 32 / 128 (ragged words and tiles)."""
    from oracle import mc_oracle
    from sculptmate_b200 import runtime

    g = golden("extract_mesh.npz")
    m = _model(g)
    tp = torch.from_numpy(g["triplane"]).cuda()
    dens0 = m.renderer.query_lattice(m.decoder, tp, R)
    thr = float(dens0.median())
    nwords = R * R * ((R + 31) // 32)
    ws = runtime._mc_cache.get(dens0.device, (R, R, R)).ws
    dens = m.renderer.query_lattice(m.decoder, tp, R, mc_signs=(thr, 1.0))
    torch.cuda.synchronize()
    fused = ws[: 4 * nwords].clone().view(torch.int32)
    assert torch.equal(dens, dens0)
    v1, f1, _ = runtime.mc_extract(dens, sub=thr, sign=1.0, flags=7, vdiv=float(R - 1.0), vmul=2 * RADIUS, vadd=-RADIUS, presigned=True)
    ws.zero_()
    v2, f2, _ = runtime.mc_extract(dens, sub=thr, sign=1.0, flags=7, vdiv=float(R - 1.0), vmul=2 * RADIUS, vadd=-RADIUS)
    torch.cuda.synchronize()
    assert torch.equal(fused, ws[: 4 * nwords].view(torch.int32))
    assert torch.equal(v1, v2) and torch.equal(f1, f2)
    v_ref, f_ref, _ = mc_oracle.marching_cubes_slab(dens.cpu().numpy(), sub=np.float32(thr), flags=7, vdiv=float(R - 1.0), vmul=2 * RADIUS, vadd=-RADIUS)
    np.testing.assert_array_equal(v1.cpu().numpy(), v_ref)
    np.testing.assert_array_equal(f1.cpu().numpy(), f_ref)


def test_threshold_errors_like_reference(golden):
    g = golden("extract_mesh.npz")
    m = _model(g)
    tp = torch.from_numpy(g["triplane"]).cuda()
    with pytest.raises(ValueError, match="within volume data range"):
        m.extract_mesh(tp[None], resolution=16, threshold=25.0)  # the reference default on random-init density


def test_batch_of_scene_codes(golden):
    g = golden("extract_mesh.npz")
    m = _model(g)
    tp = torch.from_numpy(g["triplane"]).cuda()
    codes = torch.stack([tp, tp.flip(-1), tp])
    m.extract_mesh(codes, resolution=24, threshold=float(g["threshold"]), mesh_name="b")
    assert len(m.meshes) == 3
    np.testing.assert_array_equal(m.meshes[0][1], m.meshes[2][1])
    assert not np.array_equal(m.meshes[0][0].shape, ()) and m.meshes[1][0].shape != ()


def test_capi_host_extractor_matches_python_path(golden):
    from sculptmate_b200 import _capi

    lib = _capi.load()
    g = golden("extract_mesh.npz")
    ws, bs = golden_decoder(g)
    fpp = ctypes.POINTER(ctypes.c_float)
    ws_c = [np.ascontiguousarray(w) for w in ws]
    bs_c = [np.ascontiguousarray(b) for b in bs]
    W = (fpp * 10)(*[w.ctypes.data_as(fpp) for w in ws_c])
    B = (fpp * 10)(*[b.ctypes.data_as(fpp) for b in bs_c])
    ex = ctypes.c_void_p()
    assert lib.smb_extractor_create(W, B, 9, RADIUS, -1.0, 16, 16, ctypes.byref(ex)) == 0
    try:
        tp = np.ascontiguousarray(g["triplane"])
        vp, fp_ = fpp(), ctypes.POINTER(ctypes.c_int64)()
        nv, nt = ctypes.c_int64(), ctypes.c_int64()
        R, thr = int(g["resolution"]), float(g["threshold"])
        rc = lib.smb_extract_mesh_host(ex, tp.ctypes.data_as(fpp), R, thr, ctypes.byref(vp), ctypes.byref(fp_), ctypes.byref(nv), ctypes.byref(nt))
        assert rc == 0
        v = np.ctypeslib.as_array(vp, shape=(nv.value, 3)).copy()
        f = np.ctypeslib.as_array(fp_, shape=(nt.value, 3)).copy()
        m = _model(g)
        v2, f2 = m.extract_mesh_tensors(torch.from_numpy(tp).cuda(), R, thr)
        # the C host restates aten's linspace exactly (smb_lattice_axis_host), so without any help from torch the
        # C path is bit-identical to the Python drop-in
        np.testing.assert_array_equal(f, f2.cpu().numpy())
        np.testing.assert_array_equal(v, v2.cpu().numpy())
        # and stays so when the host supplies torch-built coordinates
        from sculptmate_b200 import runtime

        axis = np.ascontiguousarray(runtime.lattice_axis(R, RADIUS).numpy())
        assert lib.smb_extractor_set_axis(ex, R, axis.ctypes.data_as(fpp)) == 0
        rc = lib.smb_extract_mesh_host(ex, tp.ctypes.data_as(fpp), R, thr, ctypes.byref(vp), ctypes.byref(fp_), ctypes.byref(nv), ctypes.byref(nt))
        assert rc == 0
        v = np.ctypeslib.as_array(vp, shape=(nv.value, 3)).copy()
        f = np.ctypeslib.as_array(fp_, shape=(nt.value, 3)).copy()
        assert np.array_equal(f, f2.cpu().numpy()) and np.array_equal(v, v2.cpu().numpy())
        # a larger lattice: the first call takes the single-pass path, the following ones the slab pipeline
        # (x-slabs, each copied to the host while the next is computed); all must be the same mesh
        Rb = 72
        axis = np.ascontiguousarray(runtime.lattice_axis(Rb, RADIUS).numpy())
        assert lib.smb_extractor_set_axis(ex, Rb, axis.ctypes.data_as(fpp)) == 0
        vb, fb = m.extract_mesh_tensors(torch.from_numpy(tp).cuda(), Rb, thr)
        for _ in range(3):
            rc = lib.smb_extract_mesh_host(ex, tp.ctypes.data_as(fpp), Rb, thr, ctypes.byref(vp), ctypes.byref(fp_), ctypes.byref(nv), ctypes.byref(nt))
            assert rc == 0
            v = np.ctypeslib.as_array(vp, shape=(nv.value, 3)).copy()
            f = np.ctypeslib.as_array(fp_, shape=(nt.value, 3)).copy()
            assert np.array_equal(f, fb.cpu().numpy()) and np.array_equal(v, vb.cpu().numpy())
        rc = lib.smb_extract_mesh_host(ex, tp.ctypes.data_as(fpp), R, 1e9, ctypes.byref(vp), ctypes.byref(fp_), ctypes.byref(nv), ctypes.byref(nt))
        assert rc == _capi.ERR_LEVEL_RANGE
    finally:
        lib.smb_extractor_destroy(ex)


def test_capi_host_extractor_textured(golden):
    """enable_texture through the C ABI: vertex colours = the Python drop-in's colour query at the same vertices,
    loop colours = their per-loop gather with alpha 1 (what import_obj_blender assigns, system.py:133-146)."""
    from sculptmate_b200 import _capi, runtime

    lib = _capi.load()
    g = golden("extract_mesh.npz")
    ws, bs = golden_decoder(g)
    fpp = ctypes.POINTER(ctypes.c_float)
    ws_c = [np.ascontiguousarray(w) for w in ws]
    bs_c = [np.ascontiguousarray(b) for b in bs]
    W = (fpp * 10)(*[w.ctypes.data_as(fpp) for w in ws_c])
    B = (fpp * 10)(*[b.ctypes.data_as(fpp) for b in bs_c])
    ex = ctypes.c_void_p()
    assert lib.smb_extractor_create(W, B, 9, RADIUS, -1.0, 16, 16, ctypes.byref(ex)) == 0
    try:
        tp = np.ascontiguousarray(g["triplane"])
        R, thr = 48, float(g["threshold"])
        axis = np.ascontiguousarray(runtime.lattice_axis(R, RADIUS).numpy())
        assert lib.smb_extractor_set_axis(ex, R, axis.ctypes.data_as(fpp)) == 0
        m = _model(g)
        tpd = torch.from_numpy(tp).cuda()
        v2, f2 = m.extract_mesh_tensors(tpd, R, thr)
        c2 = m.renderer.query_triplane(m.decoder, v2, tpd, precision="tc")["color"].cpu().numpy()
        # the triplane may be written straight into the handle's pinned staging buffer (no pageable->pinned copy)
        pin = fpp()
        assert lib.smb_extractor_pinned_input(ex, ctypes.byref(pin)) == 0
        np.copyto(np.ctypeslib.as_array(pin, shape=tp.shape), tp)
        vp, fp_ = fpp(), ctypes.POINTER(ctypes.c_int64)()
        nv, nt = ctypes.c_int64(), ctypes.c_int64()
        assert lib.smb_extract_mesh_host(ex, pin, R, thr, ctypes.byref(vp), ctypes.byref(fp_), ctypes.byref(nv), ctypes.byref(nt)) == 0
        assert np.array_equal(np.ctypeslib.as_array(fp_, shape=(nt.value, 3)), f2.cpu().numpy())
        assert np.array_equal(np.ctypeslib.as_array(vp, shape=(nv.value, 3)), v2.cpu().numpy())
        for want_loops in (False, True, True):  # single pass or slab pipeline, with and without loop colours
            vp, fp_, cp, lp = fpp(), ctypes.POINTER(ctypes.c_int64)(), fpp(), fpp()
            nv, nt = ctypes.c_int64(), ctypes.c_int64()
            rc = lib.smb_extract_mesh_host_textured(ex, tp.ctypes.data_as(fpp), R, thr, ctypes.byref(vp), ctypes.byref(fp_), ctypes.byref(cp),
                                                    ctypes.byref(lp) if want_loops else None, ctypes.byref(nv), ctypes.byref(nt))
            assert rc == 0
            v = np.ctypeslib.as_array(vp, shape=(nv.value, 3)).copy()
            f = np.ctypeslib.as_array(fp_, shape=(nt.value, 3)).copy()
            c = np.ctypeslib.as_array(cp, shape=(nv.value, 3)).copy()
            assert np.array_equal(f, f2.cpu().numpy()) and np.array_equal(v, v2.cpu().numpy())
            np.testing.assert_array_equal(c, c2)
            assert c.min() >= 0.0 and c.max() <= 1.0
            if want_loops:
                lc = np.ctypeslib.as_array(lp, shape=(3 * nt.value, 4)).copy()
                np.testing.assert_array_equal(lc[:, :3], c[f.reshape(-1)])
                assert np.all(lc[:, 3] == 1.0)
    finally:
        lib.smb_extractor_destroy(ex)


def test_full_size_256_mesh_properties(golden):
    """256^3 (BASELINE configs[1]): closedness away from the boundary is not guaranteed for a
    field that crosses the lattice border, so check size-independent properties: determinism,
    index range, every vertex referenced, vertices on lattice edges, and oracle bit-equality
    on a 24-plane sub-slab of the same density grid."""
    from oracle import mc_oracle
    from sculptmate_b200 import runtime

    g = golden("extract_mesh.npz")
    m = _model(g)
    R = 256
    tp = torch.from_numpy(g["triplane"]).cuda()
    dens = m.renderer.query_lattice(m.decoder, tp, R)
    thr = float(dens.median())
    v, f = m.extract_mesh_tensors(tp, R, thr)
    v2, f2 = m.extract_mesh_tensors(tp, R, thr)
    assert torch.equal(v, v2) and torch.equal(f, f2)
    assert int(f.min()) == 0 and int(f.max()) == len(v) - 1
    assert len(torch.unique(f)) == len(v)
    idx = (v + RADIUS) / (2 * RADIUS) * (R - 1)
    offgrid = ((idx - idx.round()).abs() > 1e-3).sum(dim=1)
    assert int(offgrid.max()) <= 1
    sl = dens[120:144].contiguous()
    pend = runtime.mc_count(sl, sub=thr, sign=1.0, emit_last_plane=False)
    vs, fs = runtime.mc_emit(pend, x_origin=120, flags=7, vdiv=float(R - 1.0), vmul=1.74, vadd=-0.87)
    vr, fr, _ = mc_oracle.marching_cubes_slab(sl.cpu().numpy(), sub=np.float32(thr), x_origin=120, emit_last_plane=False, flags=7,
                                              vdiv=float(R - 1.0), vmul=1.74, vadd=-0.87)
    np.testing.assert_array_equal(vs.cpu().numpy(), vr)
    np.testing.assert_array_equal(fs.cpu().numpy(), fr)


@pytest.mark.parametrize("R", [256, 512])
def test_full_grid_mesh_bit_exact_vs_oracle(golden, R):
    """BASELINE configs[1] / [2] sizes, WHOLE grid (not sub-slabs): the mesh of the public device path
    (TSR.extract_mesh_tensors: fused sign ballot, speculative emit) equals the C oracle's on the same density grid,
    vertices and faces bit for bit; int32 faces carry the same indices."""
    from oracle import mc_oracle

    g = golden("extract_mesh.npz")
    m = _model(g)
    tp = torch.from_numpy(g["triplane"]).cuda()
    ex_dens = m.renderer.query_lattice(m.decoder, tp, R)
    thr = float(ex_dens[:: max(1, R // 128)].median())
    v, f = m.extract_mesh_tensors(tp, R, thr)
    v32, f32 = m.extract_mesh_tensors(tp, R, thr, faces_dtype=torch.int32)
    assert f32.dtype == torch.int32 and torch.equal(v, v32) and torch.equal(f, f32.to(torch.int64))
    vr, fr, _ = mc_oracle.marching_cubes_slab(ex_dens.cpu().numpy(), sub=np.float32(thr), flags=7, vdiv=float(R - 1.0), vmul=2 * RADIUS, vadd=-RADIUS)
    assert len(vr) > 100000
    np.testing.assert_array_equal(v.cpu().numpy(), vr)
    np.testing.assert_array_equal(f.cpu().numpy(), fr)


def test_scene_cache_is_not_fooled_by_a_recycled_address(golden):
    """The add-on's loop: scene_codes = model(img); extract_mesh(scene_codes); next image.  The freed scene code's
    block is handed to the next one by the caching allocator (same address, same shape, _version 0): the renderer must
    prepare the NEW planes, not reuse the previous image's (round-1 bug: cache keyed on data_ptr)."""
    g = golden("extract_mesh.npz")
    m = _model(g)
    R = 48
    base = torch.from_numpy(g["triplane"])
    tp = base.clone().cuda()
    ptr = tp.data_ptr()
    d1 = m.renderer.query_lattice(m.decoder, tp, R).clone()
    del tp
    hits = 0
    for k in range(4):
        tp2 = (base * (1.0 + 0.25 * (k + 1))).cuda()  # a different scene, very likely in the freed block
        hits += int(tp2.data_ptr() == ptr)
        d2 = m.renderer.query_lattice(m.decoder, tp2, R)
        ref = m.renderer.query_lattice(m.decoder, tp2.clone(), R)
        assert torch.equal(d2, ref) and not torch.equal(d2, d1)
        c = m.renderer.query_triplane(m.decoder, torch.zeros(5, 3, device="cuda"), tp2)["density"]
        c_ref = m.renderer.query_triplane(m.decoder, torch.zeros(5, 3, device="cuda"), tp2.clone())["density"]
        assert torch.equal(c, c_ref)
        del tp2
    assert hits > 0, "the allocator never reused the block: the test did not exercise the aliasing case"


def test_pending_emit_survives_another_count_of_the_same_shape():
    """count(A), count(B), emit(A) with equal shapes: the workspace is shared per shape, so emit(A) must notice that
    its records were overwritten and redo A's pass instead of emitting B's records."""
    from conftest import volume
    from sculptmate_b200 import runtime

    a = torch.from_numpy(volume("gyroid", 40)).cuda()
    b = torch.from_numpy(volume("torus", 40)).cuda()
    pa = runtime.mc_count(a, sub=0.0, sign=1.0)
    va, fa = runtime.mc_emit(pa, flags=3, vdiv=39.0)
    pa2 = runtime.mc_count(a, sub=0.0, sign=1.0)
    pb = runtime.mc_count(b, sub=0.0, sign=1.0)
    assert pa2.stale() and not pb.stale()
    va2, fa2 = runtime.mc_emit(pa2, flags=3, vdiv=39.0)
    assert torch.equal(va, va2) and torch.equal(fa, fa2)
    vb, fb = runtime.mc_emit(runtime.mc_count(b, sub=0.0, sign=1.0), flags=3, vdiv=39.0)
    assert (vb.shape, fb.shape) != (va.shape, fa.shape)
