"""CPU tests of the C-ABI boundary: the library loads without a GPU, exports every symbol the header
declares, the ctypes mirror matches the header's struct layout, and argument validation returns the
documented error codes before any CUDA call."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "goi_raster.h")


def _lib():
    from diff_gaussian_rasterization import _C
    return _C, _C.lib()


def test_library_exports_every_declared_symbol():
    _C, L = _lib()
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = set(re.findall(r"\b(goi_[a-z_0-9]+)\s*\(", src))
    declared -= {"goi_alloc_fn"}
    assert {"goi_forward", "goi_backward", "goi_forward_prepare", "goi_forward_render", "goi_trace",
            "goi_mark_visible", "goi_mask", "goi_geom_bytes", "goi_last_error"} <= declared
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/goi_raster.h but not exported"
        assert name in _C.SYMBOLS, f"{name} has no ctypes prototype"
    assert L.goi_abi_version() == _C.GOI_ABI_VERSION


def test_ctypes_structs_match_header_layout(tmp_path):
    _C, _ = _lib()
    names = ["goi_view", "goi_gaussians", "goi_fwd_out", "goi_bwd_in", "goi_bwd_out", "goi_mask_args", "goi_stats"]
    prog = tmp_path / "sz.c"
    body = "".join(f'printf("{n} %zu\\n", sizeof({n}));\n' for n in names)
    body += 'printf("off_view_background %zu\\n", offsetof(goi_view, background));\n'
    body += 'printf("off_gauss_means3D %zu\\n", offsetof(goi_gaussians, means3D));\n'
    body += 'printf("off_mask_x %zu\\n", offsetof(goi_mask_args, x));\n'
    body += 'printf("off_mask_sim %zu\\n", offsetof(goi_mask_args, sim));\n'
    prog.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "{HEADER}"\nint main(){{\n{body}return 0;}}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", str(prog), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for n in names:
        assert C.sizeof(getattr(_C, n)) == int(out[n]), n
    assert _C.goi_view.background.offset == int(out["off_view_background"])
    assert _C.goi_gaussians.means3D.offset == int(out["off_gauss_means3D"])
    assert _C.goi_mask_args.x.offset == int(out["off_mask_x"])
    assert _C.goi_mask_args.sim.offset == int(out["off_mask_sim"])


def test_scratch_sizes_are_monotone_and_host_only():
    _, L = _lib()
    assert L.goi_geom_bytes(0, 0) > 0
    a, b = L.goi_geom_bytes(1000, 16), L.goi_geom_bytes(1_000_000, 16)
    assert 0 < a < b and b > 1_000_000 * 80
    assert L.goi_binning_bytes(0) > 0 and L.goi_binning_bytes(4_000_000) >= 4_000_000 * 16
    assert L.goi_image_bytes(1600, 1000) >= 1600 * 1000 * 4 + 6300 * 8


def test_validation_errors_without_gpu():
    _C, L = _lib()
    view = _C.goi_view(64, 64, 0.5, 0.5, 1.0, 3, 0, 0, 16, 16, 16, 16)     # fake non-NULL device pointers
    R = C.c_int64(-1)
    # neither SHs nor precomputed colours
    g = _C.goi_gaussians(10, 0, 0, 0, 16, None, None, None, 16, 16, 16, None)
    rc = L.goi_forward_prepare(C.byref(view), C.byref(g), 16, 16, 1 << 20, None, C.byref(R))
    assert rc == -1 and b"excatly one of either SHs or precomputed colors" in L.goi_last_error()
    # both scale/rotation and precomputed covariance
    g = _C.goi_gaussians(10, 16, 0, 0, 16, 16, None, None, 16, 16, 16, 16)
    rc = L.goi_forward_prepare(C.byref(view), C.byref(g), 16, 16, 1 << 20, None, C.byref(R))
    assert rc == -1 and b"exactly one of either scale/rotation pair" in L.goi_last_error()
    # too many channels
    g = _C.goi_gaussians(10, 16, 65, 0, 16, 16, None, 16, 16, 16, 16, None)
    rc = L.goi_forward_prepare(C.byref(view), C.byref(g), 16, 16, 1 << 20, None, C.byref(R))
    assert rc == -4 and b"GOI_MAX_SEM" in L.goi_last_error()
    # workspace too small
    g = _C.goi_gaussians(10, 16, 0, 0, 16, 16, None, None, 16, 16, 16, None)
    rc = L.goi_forward_prepare(C.byref(view), C.byref(g), 16, 16, 8, None, C.byref(R))
    assert rc == -3
    # SH degree needs more coefficients than given
    g = _C.goi_gaussians(10, 4, 0, 0, 16, 16, None, None, 16, 16, 16, None)
    rc = L.goi_forward_prepare(C.byref(view), C.byref(g), 16, 16, 1 << 20, None, C.byref(R))
    assert rc == -1 and b"sh_degree" in L.goi_last_error()
    # P == 0 is a valid no-op
    g = _C.goi_gaussians(0, 0, 0, 0, None, None, None, None, None, None, None, None)
    rc = L.goi_forward_prepare(C.byref(view), C.byref(g), None, None, 0, None, C.byref(R))
    assert rc == 0 and R.value == 0
    # goi_forward_async: argument checks come before any CUDA call
    g = _C.goi_gaussians(10, 16, 0, 0, 16, 16, None, None, 16, 16, 16, None)
    out = _C.goi_fwd_out(16, None, 16, 16, 16)
    status = (C.c_uint32 * 4)()
    rc = L.goi_forward_async(C.byref(view), C.byref(g), C.byref(out), 16, 1 << 20, 16, 1 << 20, 0, 16, 1 << 20, None, status)
    assert rc == -1 and b"capacity" in L.goi_last_error()
    rc = L.goi_forward_async(C.byref(view), C.byref(g), C.byref(out), 16, 1 << 20, 16, 1 << 20, 1000, 16, 1 << 20, None, None)
    assert rc == -1 and b"status_host" in L.goi_last_error()
    rc = L.goi_forward_async(C.byref(view), C.byref(g), C.byref(out), 16, 1 << 20, 16, 64, 1_000_000, 16, 1 << 20, None, status)
    assert rc == -3 and b"binning" in L.goi_last_error()
    # mask argument checks
    a = _C.goi_mask_args()
    a.N, a.S, a.K, a.D = 10, 0, 300, 256
    assert L.goi_mask(C.byref(a), None) == -1


def test_missing_library_is_a_hard_error():
    code = ("import sys; sys.path.insert(0, %r); import diff_gaussian_rasterization as d; d._C.lib()"
            % os.path.join(ROOT, "goi-hyperplane_b200"))
    env = dict(os.environ, GOI_RASTER_LIB="/nonexistent/libgoi_raster.so")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU/eager fallback" in r.stderr


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "goi-hyperplane_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "goi_oracle" not in txt, f
