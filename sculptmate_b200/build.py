"""Build the C-ABI shared library ``sculptmate_b200/lib/libsculptmate_b200.so``.

Plain ``nvcc`` for sm_100a only (cross-compiles without a GPU).  The library is
built IN-TREE so that it travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from typing import List

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsculptmate_b200.so")
SOURCES = ["capi.cu", "field_f32.cu", "lattice_api.cu", "mcubes.cu", "sf3d.cu", "field_pts_tc.cu", "field_tc_ta.cu", "tetgrid_tc.cu", "mesh_io.cu", "render.cu", "bake.cu"]
# developer build (`python -m sculptmate_b200.build --dev` or SMB_DEV_VARIANTS=1): superseded / experimental lattice kernels and their
# instrumentation, compiled with -DSMB_DEV_VARIANTS; the product library does not contain them
DEV_SOURCES = ["field_tc.cu", "field_tc_pair.cu"]
HEADERS = ["field_common.cuh", "field_tc_common.cuh", "ptx_sm100.cuh", "mc_tables.h", os.path.join("..", "..", "include", "sculptmate_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    # no --use_fast_math: IEEE division / expf are part of the parity contract
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built")


def _dev() -> bool:
    return bool(os.environ.get("SMB_DEV_VARIANTS")) or "--dev" in sys.argv


def _fingerprint() -> str:
    h = hashlib.sha256()
    h.update(b"dev" if _dev() else b"product")
    for name in SOURCES + DEV_SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "build.stamp")
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == fp:
                return LIB_PATH
    nvcc = _nvcc()
    objs: List[str] = []
    procs = []
    dev = _dev()
    for src in SOURCES + (DEV_SOURCES if dev else []):
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *(["-DSMB_DEV_VARIANTS"] if dev else []), "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(LIB_DIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose or failed:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see sculptmate_b200/lib/build.log")
    subprocess.check_call([nvcc, "-Wno-deprecated-gpu-targets", "-shared", "-o", LIB_PATH, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    with open(stamp, "w") as f:
        f.write(fp)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
