"""Build libdm_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library has no torch / pybind dependency: plain `nvcc -shared`.  The built `.so` is git-ignored but
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libdm_b200.so"
SOURCES = ["dm_api.cu", "dm_gemm.cu", "dm_elem.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libdm_b200.so")


HASH_PATH = LIB_DIR / "libdm_b200.so.srchash"


def _source_hash() -> str:
    """Digest of every source the library is built from (+ the flags).  Content, not mtimes: the repo snapshot that
    travels to the GPU box does not preserve modification times."""
    import hashlib

    deps = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")))
    deps.append(PKG_DIR.parent / "include" / "dm_b200.h")
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in deps:
        h.update(d.name.encode())
        h.update(d.read_bytes())
    return h.hexdigest()


def _stale() -> bool:
    if not LIB_PATH.exists() or not HASH_PATH.exists():
        return True
    return HASH_PATH.read_text().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library. Returns its path."""
    if not force and not _stale():
        return LIB_PATH
    LIB_DIR.mkdir(exist_ok=True)
    # one builder at a time (torchrun starts every rank at once): exclusive lock, then re-check
    import fcntl

    with open(LIB_DIR / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale():
            return LIB_PATH
        return _build_locked(verbose)


def _build_locked(verbose: bool) -> Path:
    obj_dir = PKG_DIR / "build" / f"pid{os.getpid()}"
    obj_dir.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = obj_dir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out)
    tmp = LIB_PATH.with_suffix(f".so.tmp{os.getpid()}")
    cmd = [nvcc, "-shared", "-o", str(tmp), *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB_PATH)
    HASH_PATH.write_text(_source_hash())
    shutil.rmtree(obj_dir, ignore_errors=True)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
