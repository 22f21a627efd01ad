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
SOURCES = ["capi.cu", "field_f32.cu", "field_tc.cu", "mcubes.cu", "sf3d.cu", "field_pts_tc.cu", "field_tc_ta.cu", "field_tc_pair.cu", "mesh_io.cu", "render.cu", "bake.cu"]
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


def _fingerprint() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
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
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj]
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
