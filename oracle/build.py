"""Build recipes for the TEST-ONLY checkers under oracle/ (not product code).

* ``build_oracle()``  -> oracle/liboracle.so   : gcc build of goi_oracle.c (CPU restatement).
* ``build_ref()``     -> oracle/_ref/libref_S{S}.so : the reference's OWN CUDA core
  (cuda_rasterizer/{forward,backward,rasterizer_impl}.cu) compiled for sm_100a from the
  sources where they lie under /root/reference, plus oracle/ref_shim.cu (our C-ABI shim
  around CudaRasterizer::Rasterizer).  No reference source is copied: the channel count
  (a compile-time ``#define SEM_CHANNELS 10`` in cuda_rasterizer/config.h:18) is chosen
  by pre-including oracle/ref_cfg/config_S{S}.h, which defines the reference header's own
  include guard so its config.h becomes a no-op.  Outputs go only to oracle/_ref/
  (git-ignored, but shipped to the GPU box by gpurun).  The reference's cmake/setup.py
  build system is not used.

Run ``python oracle/build.py [oracle] [ref]``.  /root/reference exists only in the
authoring container; on the GPU box the prebuilt .so files are used as-is.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference/submodules/diff-gaussian-rasterization"
REF_OUT = os.path.join(HERE, "_ref")
REF_CHANNELS = (1, 4, 8, 10, 16, 32)    # S=0 requests are served by the S=1 build fed with zeros
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "goi_oracle.c")
    out = os.path.join(HERE, "liboracle.so")
    if not force and _newer(out, [src]):
        return out
    cmd = ["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off", "-fopenmp",
           "-fno-fast-math", "-Wall", "-Wno-unknown-pragmas", src, "-o", out, "-lm"]
    subprocess.run(cmd, check=True)
    return out


def ref_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "cuda_rasterizer"))


def _build_ref_one(S: int, force: bool) -> str:
    os.makedirs(REF_OUT, exist_ok=True)
    out = os.path.join(REF_OUT, f"libref_S{S}.so")
    cr = os.path.join(REF_ROOT, "cuda_rasterizer")
    srcs = [os.path.join(cr, f) for f in ("forward.cu", "backward.cu", "rasterizer_impl.cu")]
    shim = os.path.join(HERE, "ref_shim.cu")
    cfg = os.path.join(HERE, "ref_cfg", f"config_S{S}.h")
    if not force and _newer(out, srcs + [shim, cfg]):
        return out
    cmd = [NVCC, "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-w",
           "-include", "cstdint",                 # rasterizer_impl.h uses uintptr_t without the header
           "-include", cfg,                       # overrides config.h through its include guard
           "-I", os.path.join(REF_ROOT, "third_party", "glm"),
           "-I", REF_ROOT,
           *srcs, shim, "-o", out]
    subprocess.run(cmd, check=True)
    return out


def build_ref(force: bool = False, channels=REF_CHANNELS) -> list[str]:
    if not ref_available():
        return [p for p in (os.path.join(REF_OUT, f"libref_S{S}.so") for S in channels) if os.path.exists(p)]
    with ThreadPoolExecutor(max_workers=4) as ex:
        return list(ex.map(lambda S: _build_ref_one(S, force), channels))


if __name__ == "__main__":
    what = sys.argv[1:] or ["oracle", "ref"]
    if "oracle" in what:
        print(build_oracle(force="--force" in what))
    if "ref" in what:
        for p in build_ref(force="--force" in what):
            print(p)
