"""In-tree build of the CUDA engine (libjaadb200.so) for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT_DIR = os.path.join(_HERE, "_build")
LIB = os.path.join(OUT_DIR, "libjaadb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # Java float semantics: separate IEEE multiply and add, never a fused multiply-add.
    "--fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                out.append(os.path.join(root, f))
    out.append(os.path.join(_HERE, "..", "include", "jaadb200.h"))
    return out


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libjaadb200.so")
    os.makedirs(OUT_DIR, exist_ok=True)
    extra = os.environ.get("JAADB200_NVCC_DEFS", "").split()   # tuning experiments only, e.g. -DK2_STEREO_MIN_BLOCKS=6
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "jaadb_engine.cu"),
                                                                                  os.path.join(CSRC, "container_index.cpp")]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB
