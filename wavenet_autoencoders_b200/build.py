"""Build libwae_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m wavenet_autoencoders_b200.build [--force]

The library links the CUDA runtime statically and resolves the one driver symbol it needs
(cuTensorMapEncodeTiled) at run time through cudaGetDriverEntryPoint, so it loads (and its
symbols can be inspected) on a machine without a GPU or libcuda -- computing needs a B200.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libwae_b200.so"
OBJ = PKG / "build"

SOURCES = ["wae_lib.cu", "vq_search.cu", "wn_stack_f32.cu", "wn_stack_bf16.cu", "wn_ar.cu", "wn_train.cu", "wn_bwd.cu", "synth_post.cu", "enc_vq_fused.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--cudart", "static",
]


def _extra_flags() -> list:
    flags = []
    if os.environ.get("WAE_LAYER_PROF") == "1":
        flags.append("-DWAE_LAYER_PROF")      # role counters of the residual-layer kernels (tools/layer_profile.py)
    if os.environ.get("WAE_AR_PROF") == "1":
        flags.append("-DWAE_AR_PROF")         # phase counters of the autoregressive kernels (tools/ar_profile.py)
    for d in os.environ.get("WAE_NVCC_DEFS", "").split():
        flags.append("-D" + d)                # experiment switches (e.g. WAE_V4_STAGES=3), never set in a shipped build
    return flags


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "wae_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + _extra_flags()).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link libwae_b200.so. Returns the library path."""
    stamp = OBJ / "stamp"
    dig = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    OBJ.mkdir(exist_ok=True)
    nvcc = _nvcc()
    sources = [s for s in SOURCES if (CSRC / s).exists()]

    def compile_one(src: str) -> Path:
        obj = OBJ / (src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *_extra_flags(), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    cmd = [nvcc, "-shared", "--cudart", "static", "-o", str(LIB), *map(str, objs), "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(dig)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
