"""Compile the CUDA kernels + C ABI into goi-hyperplane_b200/lib/libgoi_raster.so for sm_100a.

    python goi-hyperplane_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  Each .cu becomes an object (in parallel), then one shared
library with a static CUDA runtime, so the only run-time dependency is the driver.  The .so is
git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "lib", "libgoi_raster.so")
LIB_SEMLOSS = os.path.join(HERE, "lib", "libgoi_semloss.so")      # fused training loss (its own tcgen05 GEMMs: no library dependency)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["api.cu", "preprocess.cu", "binning.cu", "composite_fwd.cu", "composite_bwd.cu", "mask.cu"]
HEADERS = [os.path.join(CSRC, "goi_internal.cuh"), os.path.join(CSRC, "goi_cull.cuh"), os.path.join(CSRC, "semloss_tc.cuh"),
           os.path.join(HERE, "..", "include", "goi_raster.h"), os.path.join(HERE, "..", "include", "goi_semloss.h")]
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, force: bool, verbose: bool, stats: bool = False) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".stats.o" if stats else ".o"))
    path = os.path.join(CSRC, src)
    if force or _stale(obj, [path] + HEADERS):
        cmd = [NVCC, *FLAGS, *(["-DGOI_STATS"] if stats else []), "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    return obj


def build(force: bool = False, verbose: bool = False, stats: bool = False) -> str:
    """stats=True builds the instrumented twin lib/libgoi_raster_stats.so (work counters in the composite
    kernels, -DGOI_STATS) used by profiles/work_counters.py; the product library never carries them."""
    os.makedirs(OBJ, exist_ok=True)
    lib = LIB.replace(".so", "_stats.so") if stats else LIB
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose, stats), SOURCES))
    if force or _stale(lib, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
               "-Xcompiler", "-fPIC", *objs, "-o", lib]
        subprocess.run(cmd, check=True)
    if not stats:
        build_semloss(force, verbose)
    return lib


def build_semloss(force: bool = False, verbose: bool = False) -> str:
    """libgoi_semloss.so: csrc/semloss.cu + csrc/semloss_tc.cuh (static CUDA runtime; no cuBLAS -- both contractions of
    the loss are hand-written tcgen05 kernels).  Kept apart from libgoi_raster.so: a trainer-side library."""
    obj = _compile("semloss.cu", force, verbose)
    if force or _stale(LIB_SEMLOSS, [obj]):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
               "-Xcompiler", "-fPIC", obj, "-o", LIB_SEMLOSS]
        subprocess.run(cmd, check=True)
    return LIB_SEMLOSS


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, stats="--stats" in sys.argv))
