"""No-GPU checks of the C ABI: the library loads, exports every symbol the header
declares, and its host-side functions (decoder packing, lattice axis) are right."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from sculptmate_b200 import _capi, build

    build.build()
    return _capi.load()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "sculptmate_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from sculptmate_b200 import _capi

    names = _header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sculptmate_b200.h but not exported"
        assert n in _capi.SIGNATURES, f"{n} has no ctypes signature in _capi.py"
    assert sorted(_capi.SIGNATURES) == names


def test_status_strings_and_device_check(lib):
    from sculptmate_b200 import _capi

    assert _capi.status_string(0) == "ok"
    assert "volume data range" in _capi.status_string(_capi.ERR_LEVEL_RANGE)
    assert "No surface" in _capi.status_string(_capi.ERR_NO_SURFACE)
    if not torch.cuda.is_available():
        assert lib.smb_device_check() != 0  # fails loudly without a B200, never falls back


def _sw128(row, k):
    return (row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) & 7) ^ (row & 7)) << 4) + (k & 7) * 2


def test_decoder_pack_layout_and_values(lib, golden):
    from conftest import golden_decoder
    from sculptmate_b200 import runtime

    g = golden("field_small.npz")
    ws, bs = golden_decoder(g)
    blob, lay = runtime.pack_decoder_host([torch.from_numpy(w) for w in ws], [torch.from_numpy(b) for b in bs])
    b = blob.numpy()
    assert lay.n_hidden == 9 and lay.total_bytes == b.size
    # contiguity contract of the tensor-core section
    assert lay.off_tc_final == lay.off_tc_hidden + 8 * 8192
    assert lay.off_bias_half == lay.off_tc_final + 2048
    assert lay.off_bias_final == lay.off_bias_half + 9 * 64 * 4
    assert lay.off_tc_hidden % 1024 == 0 and lay.off_tc_l0 % 1024 == 0
    rng = np.random.RandomState(0)
    for _ in range(200):
        l, n, k = rng.randint(1, 9), rng.randint(64), rng.randint(64)
        off = lay.off_tc_hidden + (l - 1) * 8192 + _sw128(n, k)
        got = b[off : off + 2].view(np.float16)[0]
        assert got == np.float16(np.float32(0.5) * ws[l][n, k])
    for n in range(16):
        for k in (0, 7, 8, 63):
            off = lay.off_tc_final + _sw128(n, k)
            got = b[off : off + 2].view(np.float16)[0]
            assert got == (np.float16(ws[9][n, k]) if n < 4 else np.float16(0))
    for _ in range(100):
        n, k = rng.randint(64), rng.randint(128)
        off = lay.off_tc_l0 + (k // 64) * 8192 + _sw128(n, k % 64)
        got = b[off : off + 2].view(np.float16)[0]
        assert got == (np.float16(np.float32(0.5) * ws[0][n, k]) if k < 120 else np.float16(0))
    # bias K-blocks: K rows 0/1 = fp16 hi/lo of b_l/2, everything else zero
    assert lay.off_tc_biasblk % 1024 == 0 and lay.off_tc_biasblk + 8 * 8192 <= lay.total_bytes
    for l in range(1, 9):
        img = b[lay.off_tc_biasblk + (l - 1) * 8192 : lay.off_tc_biasblk + l * 8192]
        total = 0.0
        for n in range(64):
            hi = img[_sw128(n, 0) : _sw128(n, 0) + 2].view(np.float16)[0]
            lo = img[_sw128(n, 1) : _sw128(n, 1) + 2].view(np.float16)[0]
            half_b = np.float32(0.5) * bs[l][n]
            assert hi == np.float16(half_b)
            assert abs(np.float32(hi) + np.float32(lo) - half_b) <= max(2.0 ** -21 * abs(half_b), 2.0 ** -25)  # lo may be an fp16 subnormal
            total += abs(float(hi)) + abs(float(lo))
        assert np.abs(img.view(np.float16).astype(np.float64)).sum() == pytest.approx(total)
    bh = b[lay.off_bias_half : lay.off_bias_half + 9 * 64 * 4].view(np.float32).reshape(9, 64)
    np.testing.assert_array_equal(bh, np.stack([0.5 * x for x in bs[:9]]))
    np.testing.assert_array_equal(b[lay.off_bias_final : lay.off_bias_final + 16].view(np.float32), bs[9])
    w0h = b[lay.off_w0_half : lay.off_w0_half + 64 * 120 * 4].view(np.float32).reshape(64, 120)
    np.testing.assert_array_equal(w0h, 0.5 * ws[0])
    f32 = b[lay.off_f32 :].view(np.float32)
    np.testing.assert_array_equal(f32[: 64 * 120], ws[0].ravel())
    np.testing.assert_array_equal(f32[64 * 120 : 64 * 120 + 64], bs[0])


def test_decoder_pack_rejects_other_architectures():
    from sculptmate_b200 import runtime

    ws = [torch.zeros(32, 120)] + [torch.zeros(32, 32)] * 2 + [torch.zeros(4, 32)]
    bs = [torch.zeros(32)] * 3 + [torch.zeros(4)]
    with pytest.raises(NotImplementedError):
        runtime.pack_decoder_host(ws, bs)


def test_lattice_axis_host_equals_torch(lib):
    """smb_lattice_axis_host restates aten's linspace (fused multiply-add per element) and the two scale_tensor
    remaps bit for bit: a C host without torch gets the very coordinates the Python drop-in builds with the
    reference's own ops (isosurface.py:30-32 -> system.py:177-181 -> nerf_renderer.py:52-54)."""
    from sculptmate_b200 import runtime

    for R in list(range(2, 401)) + [512, 640, 1000, 1024]:
        out = np.empty(R, np.float32)
        assert lib.smb_lattice_axis_host(R, 0.87, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float))) == 0
        ref = runtime.lattice_axis(R, 0.87).numpy()
        assert np.array_equal(out, ref), f"R={R}: {int((out != ref).sum())} coordinates differ from torch"
        assert out[0] == -1.0 and (np.diff(out) > 0).all()


def test_bad_arguments_return_status_not_crash(lib):
    from sculptmate_b200 import _capi

    assert lib.smb_decoder_layout_for(1, ctypes.byref(_capi.DecoderLayout())) == _capi.ERR_BAD_ARG
    assert lib.smb_mc_workspace_bytes(0, 4, 4) == 0
    assert lib.smb_mc_workspace_bytes(256, 256, 256) > 256**3 // 32 * 16
    assert lib.smb_mc_count(None, 4, 4, 4, 0.0, 1.0, 1, None, 0, None, None) == _capi.ERR_BAD_ARG
    assert lib.smb_query_lattice_tc(None, None, None, None, None, 8, 0, 8, None, None, None) == _capi.ERR_BAD_ARG


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "sculptmate_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "mc_oracle" not in txt.replace("oracle/mc_oracle.c", ""), f


def test_cpu_tensors_fail_loudly(lib):
    from sculptmate_b200.tsr import TSR

    m = TSR()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.extract_mesh(torch.zeros(1, 3, 40, 64, 64), resolution=8, threshold=0.5)
