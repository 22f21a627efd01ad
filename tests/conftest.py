import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)

    return load


def golden_decoder(g):
    ws = [g[f"w{i}"] for i in range(10)]
    bs = [g[f"b{i}"] for i in range(10)]
    return ws, bs


# ---- synthetic volumes shared by CPU and GPU tests (SURVEY 8d family C) ----
def volume(kind: str, R: int, seed: int = 0) -> np.ndarray:
    a = np.linspace(-1, 1, R, dtype=np.float32)
    x, y, z = np.meshgrid(a, a, a, indexing="ij")
    if kind == "sphere":
        return (0.6 - np.sqrt(x * x + y * y + z * z)).astype(np.float32)
    if kind == "torus":
        q = np.sqrt(x * x + y * y) - 0.55
        return (0.22 - np.sqrt(q * q + z * z)).astype(np.float32)
    if kind == "gyroid":
        s = 3.0 * np.pi
        return (np.sin(s * x) * np.cos(s * y) + np.sin(s * y) * np.cos(s * z) + np.sin(s * z) * np.cos(s * x)).astype(np.float32)
    if kind == "noise":
        rng = np.random.RandomState(seed)
        return rng.randn(R, R, R).astype(np.float32)
    if kind == "smooth":
        rng = np.random.RandomState(seed)
        v = rng.randn(R, R, R).astype(np.float32)
        for ax in range(3):
            v = (np.roll(v, 1, ax) + v + np.roll(v, -1, ax)) / 3
        return v.astype(np.float32)
    raise ValueError(kind)


def mesh_topology(verts: np.ndarray, faces: np.ndarray, strict: bool = True):
    """(closed, euler_characteristic, signed_volume).  closed = every directed edge is
    unique and has its reverse (oriented 2-manifold without boundary); pass
    strict=False to only require that each directed edge occurs as often as its
    reverse (watertight; fan diagonals lying in an ambiguous cell face may be shared
    by the two cells, which is a touching, not a crack)."""
    f = faces.astype(np.int64)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    n = int(verts.shape[0]) + 1
    key = e[:, 0] * n + e[:, 1]
    rkey = e[:, 1] * n + e[:, 0]
    unique = (len(np.unique(key)) == len(key)) or not strict
    closed = unique and np.array_equal(np.sort(key), np.sort(rkey))
    used = len(np.unique(f))
    chi = used - len(key) // 2 + len(f)
    v = verts.astype(np.float64)
    vol = np.einsum("ij,ij->i", v[f[:, 0]], np.cross(v[f[:, 1]], v[f[:, 2]])).sum() / 6.0
    return closed, chi, vol
