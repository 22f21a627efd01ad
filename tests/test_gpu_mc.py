"""GPU parity: CUDA marching cubes vs the oracle, bit-exact, through the C ABI."""
import numpy as np
import pytest
import torch

from conftest import mesh_topology, volume

pytestmark = pytest.mark.gpu

FLAGS = 7  # FLIP | DIV | AFFINE


def _gpu_mc(g, sub, sign, emit_last=True, x_origin=0, off=0, R=None):
    from sculptmate_b200 import runtime

    R = R or max(g.shape)
    gd = torch.from_numpy(g).cuda()
    pend = runtime.mc_count(gd, sub=sub, sign=sign, emit_last_plane=emit_last)
    v, f = runtime.mc_emit(pend, x_origin=x_origin, flags=FLAGS, vdiv=float(R - 1), vmul=1.74, vadd=-0.87, vertex_id_offset=off)
    return v.cpu().numpy(), f.cpu().numpy(), pend


def _oracle_mc(g, sub, sign, emit_last=True, x_origin=0, R=None):
    from oracle import mc_oracle

    R = R or max(g.shape)
    return mc_oracle.marching_cubes_slab(g, sub=sub, sign=sign, x_origin=x_origin, emit_last_plane=emit_last, flags=FLAGS,
                                         vdiv=float(R - 1), vmul=1.74, vadd=-0.87)


CASES = [
    ("sphere", (16, 16, 16)), ("noise", (20, 20, 20)), ("gyroid", (33, 33, 33)), ("smooth", (40, 17, 70)),
    ("noise", (9, 33, 65)), ("noise", (2, 2, 2)), ("smooth", (3, 70, 2)), ("torus", (64, 64, 64)), ("gyroid", (96, 96, 96)),
]


@pytest.mark.parametrize("kind,shape", CASES)
@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_mesh_bit_exact(kind, shape, sign):
    R = max(shape)
    g = volume(kind, R, seed=4)[: shape[0], : shape[1], : shape[2]].copy()
    v, f, pend = _gpu_mc(g, 0.03, sign)
    v_ref, f_ref, c = _oracle_mc(g, 0.03, sign)
    assert (pend.nverts, pend.ntris, pend.nverts_numbered) == (c.nverts, c.ntris, c.nverts_numbered)
    np.testing.assert_array_equal(v.view(np.uint32), v_ref.view(np.uint32))
    np.testing.assert_array_equal(f, f_ref)


@pytest.mark.parametrize("kind,shape", CASES)
def test_cube_cases_bit_exact(kind, shape):
    from oracle import mc_oracle
    from sculptmate_b200 import runtime

    if min(shape) < 2:
        pytest.skip("no cells")
    R = max(shape)
    g = volume(kind, R, seed=4)[: shape[0], : shape[1], : shape[2]].copy()
    cs = runtime.mc_cases(torch.from_numpy(g).cuda(), 0.03, 1.0).cpu().numpy()
    np.testing.assert_array_equal(cs, mc_oracle.cube_cases(g, 0.03, 1.0))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slabs_on_gpu_concatenate_to_single_gpu_mesh(world):
    from sculptmate_b200.dist import slab_partition

    R = 72
    g = volume("gyroid", R)
    v_full, f_full, _ = _gpu_mc(g, 0.0, 1.0)
    vs, fs, off = [], [], 0
    for r, (a, b) in enumerate(slab_partition(R, world)):
        v, f, pend = _gpu_mc(g[a : b + 1].copy(), 0.0, 1.0, emit_last=(r == world - 1), x_origin=a, off=off, R=R)
        vs.append(v)
        fs.append(f)
        off += pend.nverts
    np.testing.assert_array_equal(np.concatenate(vs), v_full)
    np.testing.assert_array_equal(np.concatenate(fs), f_full)


def test_full_size_256_properties():
    """BASELINE config size: checked through size-independent properties (closed oriented
    surface, Euler characteristic, volume) plus bit-exactness of a sub-slab vs the oracle."""
    R = 256
    g = volume("torus", R)
    v, f, pend = _gpu_mc(g, 0.0, 1.0)
    closed, chi, vol = mesh_topology(v, f)
    assert closed and chi == 0 and vol > 0
    exact = 2 * np.pi**2 * 0.55 * 0.22**2 * (0.87 / 1.0) ** 3  # torus volume in (-.87,.87) units
    assert abs(vol / exact - 1) < 5e-3
    v2, f2, _ = _gpu_mc(g, 0.0, 1.0)
    np.testing.assert_array_equal(v, v2)  # deterministic
    np.testing.assert_array_equal(f, f2)
    sl = g[100:133].copy()
    vs, fs, _ = _gpu_mc(sl, 0.0, 1.0, x_origin=100, R=R)
    vr, fr, _ = _oracle_mc(sl, 0.0, 1.0, x_origin=100, R=R)
    np.testing.assert_array_equal(vs, vr)
    np.testing.assert_array_equal(fs, fr)


def test_helper_forward_dropin(golden):
    """MarchingCubeHelper.forward == the reference wrapper run on the oracle MC (golden)."""
    from sculptmate_b200.tsr import MarchingCubeHelper

    g = golden("helper_sphere.npz")
    h = MarchingCubeHelper(int(g["resolution"]))
    v, f = h(torch.from_numpy(g["level_in"]).cuda())
    assert v.dtype == torch.float32 and f.dtype == torch.int64 and v.is_cuda and f.is_cuda
    np.testing.assert_array_equal(v.cpu().numpy(), g["v_pos"])
    np.testing.assert_array_equal(f.cpu().numpy(), g["t_pos_idx"])
    v1, f1 = h(torch.from_numpy(g["level_in"]).cuda().view(-1))  # (R^3,) accepted like (R^3,1)
    assert torch.equal(v1, v) and torch.equal(f1, f)


def test_error_behaviour_matches_skimage():
    from sculptmate_b200.tsr import MarchingCubeHelper

    R = 8
    h = MarchingCubeHelper(R)
    with pytest.raises(ValueError, match="within volume data range"):
        h(torch.ones(R**3, 1).cuda())
    with pytest.raises(ValueError):
        h(-torch.ones(R**3, 1).cuda())
    z = torch.ones(R**3, 1)
    z[77] = 0.0
    with pytest.raises(RuntimeError, match="No surface"):
        h(z.cuda())


def test_speculative_extract_equals_two_phase_incl_overflow():
    """mc_extract launches emit behind count into buffers sized from the previous mesh of the
    same shape; a mesh that outgrows them must be re-emitted, a smaller one sliced."""
    from sculptmate_b200 import runtime

    R = 48
    runtime._mc_caps.clear()
    a = np.linspace(-1, 1, R, dtype=np.float32)
    x, y, z = np.meshgrid(a, a, a, indexing="ij")
    rr = np.sqrt(x * x + y * y + z * z)
    for radius in (0.2, 0.9, 0.5, 0.21):  # first call (two-phase), overflow, fits, fits with slack
        g = torch.from_numpy((radius - rr).astype(np.float32)).cuda()
        v, f, pend = runtime.mc_extract(g, sub=0.0, sign=1.0, flags=3, vdiv=float(R - 1))
        p2 = runtime.mc_count(g, sub=0.0, sign=1.0)
        v2, f2 = runtime.mc_emit(p2, flags=3, vdiv=float(R - 1))
        assert (pend.nverts, pend.ntris) == (p2.nverts, p2.ntris) == (v.shape[0], f.shape[0])
        assert torch.equal(v, v2) and torch.equal(f, f2), radius


@pytest.mark.parametrize("shape", [(2, 2, 2), (3, 5, 33), (7, 4, 65), (40, 70, 31), (5, 300, 9), (3, 5, 2400), (2, 3, 4700)])
def test_ragged_and_minimum_shapes_bit_exact(shape):
    """Non-cubic slabs, rows that are not a multiple of the 32-sample word, planes larger than one
    1024-word chunk, the 2x2x2 minimum, and rows longer than the emit kernel's staged halo (nz > 2272: neighbour words
    read from global memory; more than 256 words per row) -- all bit-exact against the oracle."""
    rng = np.random.RandomState(sum(shape))
    g = rng.randn(*shape).astype(np.float32)
    for a in range(3):  # a little smoothing so that surfaces are not pure noise
        g = (g + np.roll(g, 1, a)) * 0.5
    v_ref, f_ref, _ = _oracle_mc(g, np.float32(0.05), 1.0)
    v, f, pend = _gpu_mc(g, 0.05, 1.0)
    assert pend.nverts == len(v_ref) and pend.ntris == len(f_ref)
    np.testing.assert_array_equal(v, v_ref)
    np.testing.assert_array_equal(f, f_ref)


@pytest.mark.parametrize("kind,shape", [("gyroid", (33, 32, 64)), ("noise", (9, 32, 32)), ("torus", (64, 64, 64)), ("gyroid", (40, 96, 96)),
                                        ("gyroid", (37, 45, 50)), ("noise", (8, 5, 2400))])
@pytest.mark.parametrize("faces_dtype", [torch.int64, torch.int32])
def test_emit_variants_equal_the_default(kind, shape, faces_dtype):
    """The coalescing emit (MC_COALESCE: a warp's output run assembled in shared memory and written with 128-byte stores;
    what the multi-GPU gather uses for peer memory) and the int32 face width produce the same mesh as the default kernel,
    on whole grids and on an interior slab with an id offset, including ragged shapes and rows longer than the staged halo."""
    from sculptmate_b200 import _capi, runtime

    R = max(shape)
    if R <= 128:
        g = volume(kind, R, seed=3)[: shape[0], : shape[1], : shape[2]].copy()
    else:  # a long-row slab: do not build the R^3 volume
        g = np.random.RandomState(3).randn(*shape).astype(np.float32)
    gd = torch.from_numpy(g).cuda()
    for emit_last, x_origin, off in ((True, 0, 0), (False, 5, 1000)):
        pend = runtime.mc_count(gd, sub=0.01, sign=1.0, emit_last_plane=emit_last)
        v0, f0 = runtime.mc_emit(pend, x_origin=x_origin, flags=FLAGS, vdiv=float(R - 1), vmul=1.74, vadd=-0.87, vertex_id_offset=off)
        fl = FLAGS | _capi.MC_COALESCE | (_capi.MC_FACES_I32 if faces_dtype == torch.int32 else 0)
        v1, f1 = runtime.mc_emit(pend, x_origin=x_origin, flags=fl, vdiv=float(R - 1), vmul=1.74, vadd=-0.87, vertex_id_offset=off)
        assert f1.dtype == faces_dtype and pend.ntris > 0
        assert torch.equal(v0, v1) and torch.equal(f0, f1.to(torch.int64))
