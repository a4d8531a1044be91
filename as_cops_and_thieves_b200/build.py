"""In-tree build of the CUDA extension: ``nvcc`` -> ``as_cops_and_thieves_b200/libcat_b200.so``.

sm_100a only (``-gencode arch=compute_100a,code=sm_100a``), ``-lineinfo`` so ncu's source page maps
to ``csrc/*.cuh`` (``cat_b200.cu`` = C ABI + host side; ``world_kernel.cuh``, ``gae_kernels.cuh``, ``state_view.cuh`` = the kernels).  The shared library has a plain C ABI (``include/cat_b200.h``) and links
only the CUDA runtime; Python binds it with ctypes (``_lib.py``).  nvcc cross-compiles without a
GPU, so this runs on the CPU-only build box too.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
SRC = PKG / "csrc" / "cat_b200.cu"
LIB = PKG / "libcat_b200.so"
DEPS = [SRC, *sorted((PKG / "csrc").glob("*.cuh")), *sorted((PKG / "csrc").glob("*.h")), ROOT / "include" / "cat_b200.h", ROOT / "include" / "cat_philox.h"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (needed to build libcat_b200.so)")


def nvcc_cmd(out: Path = LIB, extra=()) -> list:
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-prec-div=false", "-prec-sqrt=false",
        "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-o", str(out), str(SRC),
    ]
    if Path("/usr/bin/g++").exists():
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    return cmd + list(extra)


def up_to_date() -> bool:
    return LIB.exists() and all(LIB.stat().st_mtime >= d.stat().st_mtime for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Build libcat_b200.so if it is missing or older than its sources.  Safe when several processes (one rank per GPU
    under torchrun) get here at once: the check and the build run under a file lock, nvcc writes to a temporary name and
    the finished library is renamed into place, so nobody ever loads a half-written file."""
    if up_to_date() and not force:
        return LIB
    import fcntl
    with open(PKG / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if up_to_date() and not force:          # another process built it while we waited
                return LIB
            tmp = LIB.with_name(f".{LIB.name}.{os.getpid()}.tmp")
            proc = subprocess.run(nvcc_cmd(tmp), capture_output=True, text=True)
            if verbose or proc.returncode != 0:
                sys.stderr.write(proc.stdout + proc.stderr)
            if proc.returncode != 0:
                tmp.unlink(missing_ok=True)
                raise RuntimeError("nvcc failed building libcat_b200.so")
            os.replace(tmp, LIB)
            (PKG / "build_ptxas.log").write_text(proc.stdout + proc.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
