"""Host-side logic that needs no GPU: helper semantics, config handling, partitioning,
and the multi-rank gather protocol on gloo (world_size 2 and 3) with the oracle as the
per-slab backend."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import volume
from sculptmate_b200.dist import gather_slab_meshes, slab_partition
from sculptmate_b200.tsr import TSR, MarchingCubeHelper, NeRFMLP, TriplaneNeRFRenderer
from sculptmate_b200.tsr.utils import chunk_batch, get_activation, scale_tensor


def test_scale_tensor_matches_reference_values(golden):
    g = golden("lattice.npz")
    for R in (5, 33, 64):
        ax = torch.from_numpy(g[f"axis_{R}"])
        a1 = scale_tensor(ax, (0, 1), (-0.87, 0.87))
        a2 = scale_tensor(a1, (-0.87, 0.87), (-1, 1))
        np.testing.assert_array_equal(a1.numpy(), g[f"axis_scaled_{R}"])
        np.testing.assert_array_equal(a2.numpy(), g[f"axis_unit_{R}"])


def test_grid_vertices_match_reference(golden):
    g = golden("lattice.npz")
    for R in (2, 5, 16):
        h = MarchingCubeHelper(R)
        assert h.points_range == (0, 1) and h.resolution == R
        np.testing.assert_array_equal(h.grid_vertices.numpy(), g[f"verts_{R}"])
        assert h.grid_vertices is h.grid_vertices  # cached


def test_lattice_axis_is_separable_form_of_reference_positions(golden):
    from sculptmate_b200 import runtime

    g = golden("lattice.npz")
    for R in (5, 16):
        ax = runtime.lattice_axis(R, 0.87).numpy()
        v = g[f"verts_unit_{R}"].reshape(R, R, R, 3)
        np.testing.assert_array_equal(v[:, 0, 0, 0], ax)
        np.testing.assert_array_equal(v[0, :, 0, 1], ax)
        np.testing.assert_array_equal(v[0, 0, :, 2], ax)


def test_chunk_batch_semantics():
    x = torch.arange(10.0).view(10, 1)
    f = lambda t: {"a": t * 2, "b": None}
    out = chunk_batch(f, 3, x)
    assert torch.equal(out["a"], x * 2) and out["b"] is None
    assert torch.equal(chunk_batch(lambda t: t + 1, 4, x), x + 1)
    tup = chunk_batch(lambda t: (t, t * 3), 4, x)
    assert isinstance(tup, tuple) and torch.equal(tup[1], x * 3)
    assert torch.equal(chunk_batch(lambda t: t, 0, x), x)  # 0 = no chunking
    assert chunk_batch(lambda t: None, 4, x) is None
    with pytest.raises(AssertionError):
        chunk_batch(lambda: 1, 4)
    e = chunk_batch(lambda t: t, 4, torch.zeros(0, 2))  # B == 0 still calls once
    assert e.shape == (0, 2)


def test_get_activation():
    x = torch.tensor([-1.0, 0.0, 2.0])
    assert torch.equal(get_activation("exp")(x), torch.exp(x))
    assert torch.equal(get_activation("sigmoid")(x), torch.sigmoid(x))
    assert torch.equal(get_activation(None)(x), x)
    assert torch.equal(get_activation("silu")(x), torch.nn.functional.silu(x))
    with pytest.raises(ValueError):
        get_activation("trunc_exp")  # not defined in tsr/utils.py either


def test_nerfmlp_state_dict_keys_and_eager_forward(golden):
    g = golden("field_small.npz")
    dec = NeRFMLP(dict(in_channels=120, n_neurons=64, n_hidden_layers=9, activation="silu"))
    keys = list(dec.state_dict().keys())
    assert keys == [f"layers.{i}.{n}" for i in range(0, 20, 2) for n in ("weight", "bias")]
    sd = {}
    for i in range(10):
        sd[f"layers.{2 * i}.weight"] = torch.from_numpy(g[f"w{i}"])
        sd[f"layers.{2 * i}.bias"] = torch.from_numpy(g[f"b{i}"])
    dec.load_state_dict(sd)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dec(torch.zeros(2, 5, 120))  # forward runs in the CUDA library only
    assert sum(p.numel() for p in dec.parameters()) == 41284


def test_same_seed_same_init_as_reference(golden):
    """Reference decoder under torch.manual_seed(0) (golden) == ours under the same seed."""
    g = golden("field_small.npz")
    torch.manual_seed(0)
    dec = NeRFMLP(dict(in_channels=120, n_neurons=64, n_hidden_layers=9, activation="silu"))
    for i in range(10):
        np.testing.assert_array_equal(dec.state_dict()[f"layers.{2 * i}.weight"].numpy(), g[f"w{i}"])


def test_renderer_config_and_unsupported_modes():
    r = TriplaneNeRFRenderer(dict(radius=0.87, density_activation="exp"))
    assert r.cfg.density_bias == -1.0 and r.cfg.feature_reduction == "concat"
    r.set_chunk_size(8192)
    assert r.chunk_size == 8192
    with pytest.raises(AssertionError):
        r.set_chunk_size(-1)
    with pytest.raises(ValueError):
        TriplaneNeRFRenderer({})  # radius is mandatory, as in the reference dataclass
    bad = TriplaneNeRFRenderer(dict(radius=0.87, density_activation="exp", feature_reduction="mean"))
    with pytest.raises(NotImplementedError):
        bad.query_triplane(None, torch.zeros(1, 3), torch.zeros(3, 40, 4, 4))


def test_tsr_resolution_cache_and_sink():
    m = TSR()
    m.set_marching_cubes_resolution(32)
    h = m.isosurface_helper
    m.set_marching_cubes_resolution(32)
    assert m.isosurface_helper is h
    m.set_marching_cubes_resolution(16)
    assert m.isosurface_helper is not h and m.isosurface_helper.resolution == 16
    got = []
    m.mesh_sink = lambda v, f, c, name: got.append(name)
    m.import_obj_blender(np.zeros((0, 3)), np.zeros((0, 3)), None, name="x")
    assert got == ["x"]


@pytest.mark.parametrize("R,world", [(256, 8), (512, 8), (21, 3), (9, 8), (512, 1)])
def test_slab_partition(R, world):
    parts = slab_partition(R, world)
    assert parts[0][0] == 0 and parts[-1][1] == R - 1
    assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
    sizes = [b - a for a, b in parts]
    assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
    with pytest.raises(ValueError):
        slab_partition(4, 4)


# ------------------------------------------------------------- gloo, N > 1
class _OracleSlabBackend:
    """Test double for CudaSlabBackend: same protocol, oracle arithmetic, CPU tensors."""

    device = torch.device("cpu")

    def __init__(self, grid, thr, R):
        self.grid, self.thr, self.R = grid, thr, R

    def count(self, x_begin, nx, emit_last_plane):
        from oracle import mc_oracle

        self._slab = self.grid[x_begin : x_begin + nx]
        self._last = emit_last_plane
        _, _, c = mc_oracle.marching_cubes_slab(self._slab, sub=self.thr, emit_last_plane=emit_last_plane)
        return c.nverts, c.ntris

    def emit(self, x_origin, vertex_id_offset, verts_out, faces_out):
        from oracle import mc_oracle

        v, f, _ = mc_oracle.marching_cubes_slab(
            self._slab, sub=self.thr, x_origin=x_origin, emit_last_plane=self._last,
            flags=mc_oracle.FLIP | mc_oracle.DIV, vdiv=float(self.R - 1),
        )
        verts_out.copy_(torch.from_numpy(v))
        faces_out.copy_(torch.from_numpy(f + vertex_id_offset))


def _worker(rank, world, port, R, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grid = volume("gyroid", R)
        be = _OracleSlabBackend(grid, 0.05, R)
        verts, faces, counts = gather_slab_meshes(be, R, dst=0)
        if rank == 0:
            np.savez(out_path, verts=verts.numpy(), faces=faces.numpy(), counts=np.array(counts))
        else:
            assert verts is None and faces is None
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gather_equals_single_slab(world, tmp_path):
    from oracle import mc_oracle

    R = 20
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(world, _free_port(), R, out), nprocs=world, join=True)
    got = np.load(out)
    v_ref, f_ref, _ = mc_oracle.marching_cubes_slab(
        volume("gyroid", R), sub=0.05, flags=mc_oracle.FLIP | mc_oracle.DIV, vdiv=float(R - 1)
    )
    np.testing.assert_array_equal(got["verts"], v_ref)
    np.testing.assert_array_equal(got["faces"], f_ref)
    assert got["counts"].shape == (world, 2) and got["counts"][:, 0].sum() == len(v_ref)
