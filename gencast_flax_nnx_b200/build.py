"""Builds the sm_100a shared library (C ABI of include/gencast_b200.h) in-tree with nvcc.

nvcc cross-compiles without a GPU; the resulting gencast_flax_nnx_b200/libgencast_b200.so
travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libgencast_b200.so"
STAMP_PATH = PKG_DIR / "libgencast_b200.so.stamp"

SOURCES = ["abi.cu", "gemm_tcgen05.cu", "gemm_ffma.cu", "rowwise.cu", "attention_csr.cu", "attention_tc.cu", "attention_gather.cu", "edge_fused.cu", "forward.cu", "spherical.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA kernels cannot be built")


def _sources():
    return [CSRC / s for s in SOURCES if (CSRC / s).exists()]


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "gencast_b200.h"])
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    return LIB_PATH.exists() and STAMP_PATH.exists() and STAMP_PATH.read_text().strip() == _fingerprint()


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into one shared object (objects in build/, parallel)."""
    if not force and is_current():
        return LIB_PATH
    nvcc = _nvcc()
    obj_dir = PKG_DIR.parent / "build" / "obj"
    obj_dir.mkdir(parents=True, exist_ok=True)
    procs = []
    objs = []
    for src in _sources():
        obj = obj_dir / (src.stem + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v", "-c", str(src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src.name}\n{out}")
        if p.returncode != 0:
            failed = True
    (obj_dir / "ptxas.log").write_text("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log) + "\n")
    if failed:
        raise RuntimeError("nvcc failed; see log above")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *map(str, objs)]
    subprocess.run(cmd, check=True)
    STAMP_PATH.write_text(_fingerprint())
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
