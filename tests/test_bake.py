"""SF3D texture baker: UV rasterise + barycentric interpolate against the reference's own Python
functions rasterize_cpu / interpolate_cpu (fixture tests/golden/bake.npz, oracle/make_golden_bake.py)."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("k", [0, 1])
def test_oracle_matches_reference_python_functions(golden, k):
    from oracle import bake_oracle as bo

    g = golden("bake.npz")
    rast = bo.rasterize(g[f"uv{k}"], g[f"faces{k}"], int(g[f"res{k}"]))
    np.testing.assert_array_equal(rast, g[f"rast{k}"])  # mask, triangle, barycentrics: bit-exact
    np.testing.assert_array_equal(bo.interpolate(g[f"attr{k}"], g[f"faces{k}"], rast), g[f"interp{k}"])


def test_oracle_lowest_index_wins_on_overlap():
    from oracle import bake_oracle as bo

    uv = np.array([[0.1, 0.1], [0.9, 0.1], [0.1, 0.9], [0.9, 0.9]], np.float32)
    faces = np.array([[0, 1, 2], [0, 1, 3]], np.int32)  # overlapping on purpose
    r = bo.rasterize(uv, faces, 16)
    both = (bo.rasterize(uv, faces[:1], 16)[..., 3] >= 0) & (bo.rasterize(uv, faces[1:], 16)[..., 3] >= 0)
    assert both.any() and np.all(r[both][:, 3] == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [0, 1])
def test_gpu_baker_equals_reference_golden(golden, k):
    from sculptmate_b200.sf3d import TextureBaker

    g = golden("bake.npz")
    res = int(g[f"res{k}"])
    tb = TextureBaker()
    uv, faces, attr = (torch.from_numpy(g[f"{n}{k}"]).cuda() for n in ("uv", "faces", "attr"))
    rast = tb.rasterize(uv, faces.long(), res, "cuda")  # the reference passes int64 faces and narrows them (baker.py:44)
    assert rast.shape == (res, res, 4) and rast.dtype == torch.float32
    np.testing.assert_array_equal(rast.cpu().numpy(), g[f"rast{k}"])
    np.testing.assert_array_equal(tb.get_mask(rast).cpu().numpy(), g[f"rast{k}"][..., 3] >= 0)
    out = tb.interpolate(attr, rast, faces, res, "cuda")
    np.testing.assert_array_equal(out.cpu().numpy(), g[f"interp{k}"])
    np.testing.assert_array_equal(tb(attr, uv, faces, res, "cuda").cpu().numpy(), g[f"interp{k}"])


@pytest.mark.gpu
def test_gpu_baker_vs_oracle_bake_resolution_1024():
    """Full-size bake (1024^2 texels, 20 k triangles incl. overlaps and out-of-range UVs) against the numpy oracle on a
    crop, plus size-independent properties on the whole map."""
    from oracle import bake_oracle as bo
    from sculptmate_b200.sf3d import TextureBaker

    rng = np.random.RandomState(3)
    n = 101
    gx = np.linspace(-0.05, 1.05, n)  # spills over the [0,1] texture on purpose
    u, v = np.meshgrid(gx, gx, indexing="ij")
    uv = (np.stack([u, v], -1).reshape(-1, 2) + rng.uniform(-0.003, 0.003, (n * n, 2))).astype(np.float32)
    faces = []
    for i in range(n - 1):
        for j in range(n - 1):
            a, b, c, d = i * n + j, (i + 1) * n + j, (i + 1) * n + j + 1, i * n + j + 1
            faces += [[a, b, c], [a, c, d]]
    faces = np.asarray(faces, np.int32)
    attr = rng.randn(n * n, 3).astype(np.float32)
    res = 1024
    tb = TextureBaker()
    rast = tb.rasterize(torch.from_numpy(uv).cuda(), torch.from_numpy(faces).cuda(), res, "cuda")
    out = tb.interpolate(torch.from_numpy(attr).cuda(), rast, torch.from_numpy(faces).cuda(), res, "cuda").cpu().numpy()
    r = rast.cpu().numpy()
    assert (r[..., 3] >= 0).all()  # the atlas covers the whole texture
    assert np.abs(r[..., :3].sum(-1) - 1).max() < 1e-4 and r[..., :3].min() >= 0
    tri = r[..., 3].astype(np.int64)
    np.testing.assert_array_equal(out, bo.interpolate(attr, faces, r))
    # oracle on a crop: only the triangles that can touch it
    ys, xs = slice(300, 340), slice(500, 560)
    cand = np.unique(tri[280:360, 480:580])
    sub = bo.rasterize(uv, faces[cand], res, window=(ys.start, ys.stop, xs.start, xs.stop))
    sub_tri = np.where(sub[..., 3] >= 0, cand[np.maximum(sub[..., 3].astype(np.int64), 0)], -1)
    np.testing.assert_array_equal(sub_tri, tri[ys, xs])
    np.testing.assert_array_equal(sub[..., :3], r[ys, xs, :3])
