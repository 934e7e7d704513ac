"""In-tree build of the CUDA library (nvcc, sm_100a only).  The built .so is git-ignored but travels
to the GPU box with gpurun; nothing is JIT-compiled at import time on the box unless it is missing."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libcover_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "cover_b200.cu")]
DEPENDS = SOURCES + [os.path.join(HERE, "csrc", "device_core.cuh"), os.path.join(HERE, "csrc", "area_rows.cuh"), os.path.join(HERE, "csrc", "lattice_words.cuh"), os.path.join(HERE, "csrc", "polygon_split.hpp"),
                     os.path.join(ROOT, "include", "cover_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cover_b200 has no CPU fallback and cannot be built without the CUDA toolkit")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in DEPENDS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile cover_b200/libcover_b200.so when missing or older than its sources."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH
