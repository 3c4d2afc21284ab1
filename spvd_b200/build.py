"""Build libspvd_b200.so (the C-ABI library of include/spvd_b200.h) for sm_100a with nvcc.

In-tree, no torch headers: the library only needs the CUDA runtime.  ``python -m spvd_b200.build``.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libspvd_b200.so")
SOURCES = ["error.cu", "coords.cu", "hash.cu", "pointvoxel.cu", "conv_simt.cu", "conv_umma.cu", "conv_umma_f16.cu", "conv_sp_f16.cu", "fused.cu", "runtime.cu", "plan.cu", "sampler.cu", "conv_wgrad_umma.cu", "metrics.cu", "train.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise NvccMissing("nvcc not found: libspvd_b200.so cannot be built")
    return cand


class NvccMissing(RuntimeError):
    pass


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile (if stale) and return the library path.  Safe under torchrun: ranks serialise on a lock file, the first one
    builds into temporary names and renames atomically, the others find the fresh stamp."""
    import fcntl
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.stamp")
    digest = _digest()

    def fresh():
        return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest

    if not force and fresh():
        return LIB
    with open(os.path.join(LIBDIR, "build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():           # another rank built it while we waited
                return LIB
            return _build_locked(digest, stamp, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(digest: str, stamp: str, verbose: bool) -> str:
    nvcc = _nvcc()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []

    def compile_one(src):
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    tmp = LIB + f".tmp{os.getpid()}"
    r = subprocess.run([nvcc, "-shared", "-o", tmp, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)                         # atomic: nobody ever dlopens a half-written library
    with open(stamp + ".tmp", "w") as f:
        f.write(digest)
    os.replace(stamp + ".tmp", stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
