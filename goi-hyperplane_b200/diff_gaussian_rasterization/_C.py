"""ctypes binding of libgoi_raster.so -- the stand-in for the reference's pybind11 module
``diff_gaussian_rasterization._C`` (submodules/diff-gaussian-rasterization/ext.cpp:15-20).

The four functions keep the reference's names, argument order and return tuples
(rasterize_points.h:18-96) so ``diff_gaussian_rasterization/__init__.py`` reads like the reference's:

    rasterize_gaussians(bg, means3D, colors, semantics, opacity, scales, rotations, scale_modifier,
                        cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, sh, degree,
                        campos, prefiltered, debug)
        -> (num_rendered, color, semantic, depth, alpha, radii, geomBuffer, binningBuffer, imgBuffer)
    rasterize_gaussians_backward(26 args) -> 9 gradients
    rasterize_gaussians_trace(20 args)    -> (num_rendered, color, gau_sem, num_gsem, geom, binning, img)
    mark_visible(means3D, viewmatrix, projmatrix) -> bool[P]

plus ``hyperplane_mask`` for the fused mask kernel.  PyTorch is used for device memory and the
current stream only; all compute is in the CUDA library, and a missing library is a hard error
(there is no CPU or eager fallback).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GOI_RASTER_LIB", os.path.join(_HERE, "..", "lib", "libgoi_raster.so"))
GOI_ABI_VERSION = 5
GOI_MAX_SEM = 64
GOI_MASK_APE, GOI_MASK_OSH = 0, 1
GOI_RAW_OPACITY, GOI_RAW_SCALE, GOI_RAW_ROTATION = 1, 2, 4

_f32p = C.c_void_p  # device pointers travel as integers


class goi_view(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
                ("scale_modifier", C.c_float), ("sh_degree", C.c_int32), ("prefiltered", C.c_int32),
                ("debug", C.c_int32), ("background", _f32p), ("viewmatrix", _f32p), ("projmatrix", _f32p),
                ("cam_pos", _f32p)]


class goi_gaussians(C.Structure):
    _fields_ = [("P", C.c_int32), ("M", C.c_int32), ("S", C.c_int32), ("raw_flags", C.c_int32),
                ("means3D", _f32p), ("shs", _f32p), ("colors_precomp", _f32p), ("semantics", _f32p),
                ("opacities", _f32p), ("scales", _f32p), ("rotations", _f32p), ("cov3D_precomp", _f32p),
                ("shs_rest", _f32p)]


class goi_fwd_out(C.Structure):
    _fields_ = [("out_color", _f32p), ("out_semantic", _f32p), ("out_depth", _f32p), ("out_alpha", _f32p),
                ("radii", C.c_void_p)]


class goi_bwd_in(C.Structure):
    _fields_ = [("dL_dcolor", _f32p), ("dL_dsemantic", _f32p), ("dL_ddepth", _f32p), ("dL_dalpha", _f32p),
                ("out_alpha", _f32p), ("radii", C.c_void_p)]


class goi_bwd_out(C.Structure):
    _fields_ = [("dL_dmean2D", _f32p), ("dL_dconic", _f32p), ("dL_dopacity", _f32p), ("dL_dcolor", _f32p),
                ("dL_dsemantic", _f32p), ("dL_ddepth", _f32p), ("dL_dmean3D", _f32p), ("dL_dcov3D", _f32p),
                ("dL_dsh", _f32p), ("dL_dscale", _f32p), ("dL_drot", _f32p), ("accumulate", C.c_int32),
                ("_pad", C.c_int32), ("dL_dsh_rest", _f32p)]


class goi_mask_args(C.Structure):
    _fields_ = [("N", C.c_int64), ("S", C.c_int32), ("K", C.c_int32), ("D", C.c_int32), ("mode", C.c_int32),
                ("stride_n", C.c_int64), ("stride_c", C.c_int64), ("x", _f32p), ("mlp_weight", _f32p),
                ("mlp_bias", _f32p), ("lut", _f32p), ("hyperplane_w", _f32p), ("hyperplane_b", C.c_float),
                ("log_scale", C.c_float), ("thresh", C.c_float), ("sim_table", _f32p), ("sim", _f32p),
                ("bg_mask", C.c_void_p), ("idx", C.c_void_p)]


class goi_stats(C.Structure):
    _fields_ = [("num_rendered", C.c_int64), ("num_visible", C.c_int64), ("tiles_x", C.c_int32),
                ("tiles_y", C.c_int32)]


ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)

# name -> (restype, argtypes): every symbol include/goi_raster.h declares
SYMBOLS = {
    "goi_abi_version": (C.c_int, []),
    "goi_last_error": (C.c_char_p, []),
    "goi_geom_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "goi_image_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "goi_binning_bytes": (C.c_size_t, [C.c_int64]),
    "goi_forward_prepare": (C.c_int, [C.POINTER(goi_view), C.POINTER(goi_gaussians), C.c_void_p, C.c_void_p,
                                      C.c_size_t, C.c_void_p, C.POINTER(C.c_int64)]),
    "goi_forward_render": (C.c_int, [C.POINTER(goi_view), C.POINTER(goi_gaussians), C.POINTER(goi_fwd_out),
                                     C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                     C.c_int64, C.c_void_p]),
    "goi_forward_auto": (C.c_int, [C.POINTER(goi_view), C.POINTER(goi_gaussians), C.POINTER(goi_fwd_out),
                                   C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                   C.c_void_p, C.POINTER(C.c_int64)]),
    "goi_forward_async": (C.c_int, [C.POINTER(goi_view), C.POINTER(goi_gaussians), C.POINTER(goi_fwd_out),
                                    C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p, C.c_size_t,
                                    C.c_void_p, C.c_void_p]),
    "goi_forward": (C.c_int, [C.POINTER(goi_view), C.POINTER(goi_gaussians), C.POINTER(goi_fwd_out),
                              ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p, C.c_void_p,
                              C.POINTER(C.c_int64)]),
    "goi_backward": (C.c_int, [C.POINTER(goi_view), C.POINTER(goi_gaussians), C.c_int64, C.POINTER(goi_bwd_in),
                               C.POINTER(goi_bwd_out), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "goi_trace": (C.c_int, [C.POINTER(goi_view), C.POINTER(goi_gaussians), C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_int32, ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p,
                            ALLOC_FN, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "goi_mark_visible": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "goi_mask": (C.c_int, [C.POINTER(goi_mask_args), C.c_void_p]),
    "goi_forward_mask": (C.c_int, [C.POINTER(goi_view), C.POINTER(goi_gaussians), C.POINTER(goi_fwd_out),
                                   C.POINTER(goi_mask_args), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                   C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_int64)]),
    "goi_read_stats": (C.c_int, [C.POINTER(goi_view), C.POINTER(goi_gaussians), C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.POINTER(goi_stats)]),
    "goi_timing_enable": (C.c_int, [C.c_int]),
    "goi_timing_read": (C.c_int, [C.POINTER(C.c_float)]),
    "goi_stage_name": (C.c_char_p, [C.c_int]),
    "goi_launch_count": (C.c_uint64, []),
}
GOI_NUM_STAGES = 9
# Optional gradient arena, PER DEVICE: {name: preallocated tensor} for the parameter gradients "means3D", "sh",
# "semantics", "opacities", "scales", "rotations", "colors_precomp", "cov3D_precomp".  When a device has one, the
# backward of a rasterization on THAT device writes those gradients straight into the given tensors (e.g. slices of
# one flat all-reduce buffer, see goi_b200/view_parallel.py) instead of fresh allocations -- no packing copy before
# the collective.  With accumulate=True the library ADDS each view's parameter gradients to what the arena already
# holds (goi_bwd_out.accumulate): several views per rank summed in place, one all-reduce at the end.
# Keyed by device index rather than thread-local: autograd runs backward on its own worker thread (one per device),
# so the thread that installs the arena is not the thread that consumes it.  All mutable binding state lives in
# _DeviceState objects guarded by one lock; the C library below it is re-entrant per (device, stream).
import threading


class BinningOverflow(RuntimeError):
    """An asynchronously rendered view had more instances than its binning blob could hold: its outputs (and the
    gradients back-propagated from them) are undefined.  Render the view / redo the step; the capacity estimate has
    already been raised."""


class _DeviceState:
    __slots__ = ("grad_arena", "grad_accumulate", "r_guess", "last_num_rendered", "async_binning", "pending",
                 "status_pool", "headroom")

    def __init__(self):
        self.grad_arena = None
        self.grad_accumulate = False
        self.r_guess = 0                 # running upper estimate of the instance count (sizes the binning blob)
        self.last_num_rendered = 0       # R of the most recent forward on this device
        self.async_binning = False       # goi_forward_async: no num_rendered read-back per view
        self.pending = []                # [(status tensor (pinned), cuda event, capacity)] not yet examined
        self.status_pool = []            # recycled pinned status words
        self.headroom = 1.10             # capacity = headroom x the running estimate


_state_lock = threading.Lock()
_states: dict = {}


def _dev_index(device) -> int:
    device = torch.device(device) if not isinstance(device, torch.device) else device
    if device.type != "cuda":
        return -1
    return torch.cuda.current_device() if device.index is None else device.index


def device_state(device) -> _DeviceState:
    i = _dev_index(device)
    with _state_lock:
        st = _states.get(i)
        if st is None:
            st = _states[i] = _DeviceState()
        return st


def set_grad_arena(arena, accumulate=False, device=None):
    """Install (or with arena=None remove) the gradient arena of `device` (default: the arena tensors' device, else
    the current CUDA device)."""
    if device is None:
        device = next(iter(arena.values())).device if arena else torch.device("cuda", torch.cuda.current_device())
    st = device_state(device)
    with _state_lock:
        st.grad_arena = arena
        st.grad_accumulate = bool(accumulate) and arena is not None


def set_async_binning(on: bool, device=None, headroom: float = 1.10):
    """Training-loop mode (no host sync on the fast path): forwards on `device` run through
    goi_forward_async -- the binning blob is sized for `headroom` x the running instance-count estimate and the count
    stays on the device, so the host never waits for the GPU inside a view and runs ahead of it.  Each view leaves a
    4-word status in pinned host memory; `check_async()` examines the ones that have landed and raises BinningOverflow
    if a view did not fit (call it at least once per optimisation step, with wait=True before trusting a step's
    result).  The first view on a device always takes the synchronous path to seed the estimate."""
    st = device_state(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    st.async_binning = bool(on)
    st.headroom = float(headroom)


def check_async(device=None, wait: bool = False) -> int:
    """Examine the status words of asynchronously rendered views on `device`.  wait=False: only those whose copy has
    completed; wait=True: all of them (blocks).  Returns how many views were examined; raises BinningOverflow."""
    st = device_state(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    n, worst = 0, None
    while st.pending:
        status, ev, cap = st.pending[0]
        if wait:
            ev.synchronize()
        elif not ev.query():
            break
        st.pending.pop(0)
        R, overflow, violation = int(status[0]), int(status[1]), int(status[2])
        st.status_pool.append(status)
        st.last_num_rendered = R
        st.r_guess = max(R, int(0.9 * st.r_guess))
        n += 1
        if violation:
            raise RuntimeError("Point is filtered although prefiltered is set. This shouldn't happen!")
        if overflow:
            worst = (R, cap) if worst is None or R > worst[0] else worst
    if worst is not None:
        raise BinningOverflow(f"a view produced {worst[0]} tile instances but its binning blob was sized for {worst[1]}")
    return n


def num_rendered(device=None) -> int:
    """Instance count R of the most recent forward on `device` (bench.py's roofline arithmetic)."""
    return device_state(device if device is not None else torch.device("cuda", torch.cuda.current_device())).last_num_rendered


def _grad_out(st, name, shape, f32):
    arena = st.grad_arena
    if arena is not None and name in arena:
        t = arena[name]
        if t.numel() != int(torch.Size(shape).numel()) or not t.is_contiguous():
            raise RuntimeError(f"grad arena entry {name!r} has {t.numel()} elements, expected shape {tuple(shape)}")
        return t.view(shape)
    return torch.empty(shape, **f32)


_lib = None


def lib() -> C.CDLL:
    """Load libgoi_raster.so (once).  Missing library = hard error: there is no fallback path."""
    global _lib
    if _lib is None:
        path = os.path.abspath(LIB_PATH)
        if not os.path.exists(path):
            raise ImportError(
                f"{path} not found: build it with `python goi-hyperplane_b200/build.py` "
                "(nvcc, sm_100a). diff_gaussian_rasterization has no CPU/eager fallback.")
        handle = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.goi_abi_version() != GOI_ABI_VERSION:
            raise ImportError(f"{path}: ABI {handle.goi_abi_version()} != binding {GOI_ABI_VERSION}")
        _lib = handle
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().goi_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")


def _ptr(t):
    """Device pointer of a tensor, or NULL for None / empty tensors (the reference passes the data
    pointer of `torch.Tensor([])`, i.e. nullptr, for absent optional inputs, __init__.py:286-297)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32(t, name):
    if t is None or t.numel() == 0:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _make_view(bg, viewmatrix, projmatrix, campos, W, H, tan_fovx, tan_fovy, scale_modifier, degree, prefiltered,
               debug, keep):
    bg, viewmatrix, projmatrix, campos = (_f32(bg, "bg"), _f32(viewmatrix, "viewmatrix"),
                                          _f32(projmatrix, "projmatrix"), _f32(campos, "campos"))
    keep += [bg, viewmatrix, projmatrix, campos]
    return goi_view(int(W), int(H), float(tan_fovx), float(tan_fovy), float(scale_modifier), int(degree),
                    int(bool(prefiltered)), int(bool(debug)), _ptr(bg), _ptr(viewmatrix), _ptr(projmatrix),
                    _ptr(campos))


def _make_gaussians(means3D, sh, colors, semantics, opacity, scales, rotations, cov3D_precomp, keep, raw_flags=0,
                    sh_rest=None):
    if means3D.ndim != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")      # rasterize_points.cu:58-60
    means3D = _f32(means3D, "means3D") if means3D.numel() else means3D
    sh, colors, semantics = _f32(sh, "sh"), _f32(colors, "colors_precomp"), _f32(semantics, "semantics")
    opacity, scales = _f32(opacity, "opacities"), _f32(scales, "scales")
    rotations, cov3D_precomp = _f32(rotations, "rotations"), _f32(cov3D_precomp, "cov3D_precomp")
    sh_rest = _f32(sh_rest, "sh_rest")
    keep += [means3D, sh, colors, semantics, opacity, scales, rotations, cov3D_precomp, sh_rest]
    P = means3D.shape[0]
    M = sh.shape[1] if sh is not None else 0
    if sh_rest is not None:
        if sh is None or sh.shape[1] != 1:
            raise RuntimeError("sh_rest needs sh = the DC block [P,1,3]")
        M = 1 + sh_rest.shape[1]
    S = semantics.shape[1] if semantics is not None else 0
    if semantics is not None and (semantics.ndim != 2 or semantics.shape[0] != P):
        raise RuntimeError("semantics must have dimensions (num_points, S)")
    g = goi_gaussians(P, M, S, int(raw_flags), _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(semantics), _ptr(opacity),
                      _ptr(scales), _ptr(rotations), _ptr(cov3D_precomp), _ptr(sh_rest))
    return g, means3D


def rasterize_gaussians(bg, means3D, colors, semantics, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                        prefiltered, debug, raw_flags=0, sh_rest=None):
    """RasterizeGaussiansCUDA (reference rasterize_points.cu:35-123).  raw_flags / sh_rest (not in the
    reference): the inputs are the STORED parameters and the library applies the activations (GOI_RAW_*)."""
    L = lib()
    keep = []
    dev = means3D.device
    with torch.cuda.device(dev):
        g, means3D = _make_gaussians(means3D, sh, colors, semantics, opacity, scales, rotations, cov3D_precomp, keep,
                                     raw_flags, sh_rest)
        view = _make_view(bg, viewmatrix, projmatrix, campos, image_width, image_height, tan_fovx, tan_fovy,
                          scale_modifier, degree, prefiltered, debug, keep)
        P, S, H, W = g.P, g.S, int(image_height), int(image_width)
        f32 = dict(dtype=torch.float32, device=dev)
        # the library writes every pixel / every radius: no zero fill needed (reference: torch::full x5)
        out_color = torch.empty((3, H, W), **f32)
        out_sem = torch.empty((S, H, W), **f32)
        out_depth = torch.empty((1, H, W), **f32)
        out_alpha = torch.empty((1, H, W), **f32)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        geom_bytes = L.goi_geom_bytes(P, S)
        img_bytes = L.goi_image_bytes(W, H)
        geom = torch.empty((geom_bytes,), **u8)
        img = torch.empty((img_bytes,), **u8)
        stream = _stream(dev)
        R = C.c_int64(0)
        out = goi_fwd_out(_ptr(out_color), _ptr(out_sem), _ptr(out_depth), _ptr(out_alpha), _ptr(radii))
        # The binning blob is sized BEFORE the instance count is known (a guess from recent views, with
        # headroom), so prepare + render run inside one C call and the device is refilled right after the
        # one host sync of the path; a wrong guess costs one re-allocation.
        st = device_state(dev)
        if st.async_binning and st.r_guess > 0 and P and not debug:
            # no host sync: the instance count stays on the device; the blobs are carved for `cap` instances
            cap = int(st.r_guess * st.headroom) + 4096
            bin_bytes = L.goi_binning_bytes(cap)
            binning = torch.empty((bin_bytes,), **u8)
            if not st.status_pool:                  # recycle landed status words; allocate pinned memory only in bulk
                check_async(dev, wait=False)
                if not st.status_pool:
                    if len(st.pending) >= 256:
                        check_async(dev, wait=True)
                    else:
                        st.status_pool = list(torch.zeros((64, 4), dtype=torch.int32).pin_memory().unbind(0))
            status = st.status_pool.pop()
            _check(L.goi_forward_async(C.byref(view), C.byref(g), C.byref(out), geom.data_ptr(), geom_bytes,
                                       binning.data_ptr(), bin_bytes, cap, img.data_ptr(), img_bytes, stream,
                                       status.data_ptr()), "goi_forward_async")
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            st.pending.append((status, ev, cap))
            return cap, out_color, out_sem, out_depth, out_alpha, radii, geom, binning, img
        bin_bytes = L.goi_binning_bytes(max(int(st.r_guess * 1.25), 4 * P, 1 << 16)) if P else L.goi_binning_bytes(0)
        binning = torch.empty((bin_bytes,), **u8)
        rc = L.goi_forward_auto(C.byref(view), C.byref(g), C.byref(out), geom.data_ptr(), geom_bytes,
                                binning.data_ptr(), bin_bytes, img.data_ptr(), img_bytes, stream, C.byref(R))
        if rc == -3 and R.value > 0 and L.goi_binning_bytes(R.value) > bin_bytes:
            bin_bytes = L.goi_binning_bytes(R.value)
            binning = torch.empty((bin_bytes,), **u8)
            rc = L.goi_forward_render(C.byref(view), C.byref(g), C.byref(out), geom.data_ptr(), geom_bytes,
                                      binning.data_ptr(), bin_bytes, img.data_ptr(), img_bytes, R.value, stream)
        _check(rc, "goi_forward")
        st.r_guess = max(R.value, int(0.9 * st.r_guess))
        st.last_num_rendered = int(R.value)
    return int(R.value), out_color, out_sem, out_depth, out_alpha, radii, geom, binning, img


def rasterize_gaussians_backward(bg, means3D, radii, colors, semantics, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                 dL_dout_semantic, dL_dout_depth, dL_dout_alpha, sh, degree, campos, geomBuffer, R,
                                 binningBuffer, imageBuffer, alphas, debug, raw_flags=0, sh_rest=None):
    """RasterizeGaussiansBackwardCUDA (reference rasterize_points.cu:213-306).  With sh_rest the SH gradient
    comes back as the pair (dL_dsh_dc [P,1,3], dL_dsh_rest [P,M-1,3]) in place of dL_dsh."""
    L = lib()
    keep = []
    dev = means3D.device
    with torch.cuda.device(dev):
        g, means3D = _make_gaussians(means3D, sh, colors, semantics, None, scales, rotations, cov3D_precomp, keep,
                                     raw_flags, sh_rest)
        H, W = int(alphas.shape[-2]), int(alphas.shape[-1])
        view = _make_view(bg, viewmatrix, projmatrix, campos, W, H, tan_fovx, tan_fovy, scale_modifier, degree,
                          False, debug, keep)
        P, S, M = g.P, g.S, g.M
        f32 = dict(dtype=torch.float32, device=dev)
        gc, gs_, gd, ga = (_f32(dL_dout_color, "dL_dout_color"), _f32(dL_dout_semantic, "dL_dout_semantic"),
                           _f32(dL_dout_depth, "dL_dout_depth"), _f32(dL_dout_alpha, "dL_dout_alpha"))
        alphas = _f32(alphas, "alphas")
        keep += [gc, gs_, gd, ga, alphas]
        st = device_state(dev)
        with _state_lock:
            grad_arena, grad_accumulate = st.grad_arena, st.grad_accumulate
        # fully written (or zero-initialised) inside the library: torch.empty, not torch.zeros
        dL_dmeans3D = _grad_out(st, "means3D", (P, 3), f32)
        dL_dmeans2D = torch.empty((P, 3), **f32)
        dL_dcolors = _grad_out(st, "colors_precomp", (P, 3), f32) if g.colors_precomp is not None else torch.empty((P, 3), **f32)
        dL_dsemantics = _grad_out(st, "semantics", (P, S), f32)
        dL_ddepths = torch.empty((P, 1), **f32)
        dL_dconic = torch.empty((P, 2, 2), **f32)
        dL_dopacity = _grad_out(st, "opacities", (P, 1), f32)
        has_sh, has_scale = g.shs is not None, g.scales is not None
        # in-place accumulation only when EVERY input gradient of this call lives in the arena
        need = ["means3D", "opacities"] + (["semantics"] if S else []) + (["sh"] if has_sh else ["colors_precomp"]) \
            + (["sh_rest"] if sh_rest is not None and sh_rest.numel() else []) \
            + (["scales", "rotations"] if has_scale else ["cov3D_precomp"])
        acc = int(grad_accumulate and all(n in grad_arena for n in need))
        if grad_accumulate and not acc:
            raise RuntimeError(f"gradient accumulation needs arena slots for {need}")
        dL_dcov3D = _grad_out(st, "cov3D_precomp", (P, 6), f32) if not has_scale else torch.empty((P, 6), **f32)
        split = g.shs_rest is not None
        dL_dsh_rest = None
        if split:
            dL_dsh = _grad_out(st, "sh", (P, 1, 3), f32)
            dL_dsh_rest = _grad_out(st, "sh_rest", (P, M - 1, 3), f32)
        else:
            dL_dsh = _grad_out(st, "sh", (P, M, 3), f32) if has_sh else torch.zeros((P, M, 3), **f32)
        dL_dscales = _grad_out(st, "scales", (P, 3), f32) if has_scale else torch.zeros((P, 3), **f32)
        dL_drotations = _grad_out(st, "rotations", (P, 4), f32) if has_scale else torch.zeros((P, 4), **f32)
        if P != 0:
            gin = goi_bwd_in(_ptr(gc), _ptr(gs_), _ptr(gd), _ptr(ga), _ptr(alphas), _ptr(radii))
            gout = goi_bwd_out(_ptr(dL_dmeans2D), _ptr(dL_dconic), _ptr(dL_dopacity), _ptr(dL_dcolors),
                               _ptr(dL_dsemantics), _ptr(dL_ddepths), _ptr(dL_dmeans3D), _ptr(dL_dcov3D),
                               _ptr(dL_dsh), _ptr(dL_dscales), _ptr(dL_drotations), acc, 0, _ptr(dL_dsh_rest))
            _check(L.goi_backward(C.byref(view), C.byref(g), int(R), C.byref(gin), C.byref(gout),
                                  geomBuffer.data_ptr(), _ptr(binningBuffer), imageBuffer.data_ptr(),
                                  _stream(dev)), "goi_backward")
    if split:
        dL_dsh = (dL_dsh, dL_dsh_rest)
    return (dL_dmeans2D, dL_dcolors, dL_dsemantics, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
            dL_drotations)


def _tensor_allocator(store: list, device):
    """goi_alloc_fn backed by torch uint8 tensors: the reference's resizeFunctional lambdas
    (rasterize_points.cu:27-33)."""
    def alloc(_user, nbytes):
        t = torch.empty((max(int(nbytes), 1),), dtype=torch.uint8, device=device)
        store.append(t)
        return t.data_ptr()
    return ALLOC_FN(alloc)


def rasterize_gaussians_trace(bg, means3D, colors, img_sem, opacity, scales, rotations, scale_modifier,
                              cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width,
                              sh, degree, campos, prefiltered, debug, count_per_channel=True):
    """TraceGaussiansCUDA (reference rasterize_points.cu:125-211)."""
    L = lib()
    keep = []
    dev = means3D.device
    with torch.cuda.device(dev):
        g, means3D = _make_gaussians(means3D, sh, colors, None, opacity, scales, rotations, cov3D_precomp, keep)
        img_sem = _f32(img_sem, "img_sem")
        S = img_sem.shape[0] if img_sem is not None else 0
        g.S = S
        H, W = int(image_height), int(image_width)
        view = _make_view(bg, viewmatrix, projmatrix, campos, W, H, tan_fovx, tan_fovy, scale_modifier, degree,
                          prefiltered, debug, keep)
        P = g.P
        out_color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        gau_sem = torch.empty((P, S), dtype=torch.float32, device=dev)
        num_gsem = torch.empty((P,), dtype=torch.int32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        gstore, bstore, istore = [], [], []
        ga, ba, ia = _tensor_allocator(gstore, dev), _tensor_allocator(bstore, dev), _tensor_allocator(istore, dev)
        R = C.c_int64(0)
        _check(L.goi_trace(C.byref(view), C.byref(g), _ptr(img_sem), _ptr(out_color), _ptr(gau_sem), _ptr(num_gsem),
                           _ptr(radii), int(bool(count_per_channel)), ga, None, ba, None, ia, None, _stream(dev),
                           C.byref(R)), "goi_trace")
        u8 = torch.empty((0,), dtype=torch.uint8, device=dev)
    return (int(R.value), out_color, gau_sem, num_gsem, gstore[-1] if gstore else u8,
            bstore[-1] if bstore else u8, istore[-1] if istore else u8)


def mark_visible(means3D, viewmatrix, projmatrix):
    """markVisible (reference rasterize_points.cu:308-327)."""
    L = lib()
    dev = means3D.device
    with torch.cuda.device(dev):
        m, v, p = _f32(means3D, "means3D"), _f32(viewmatrix, "viewmatrix"), _f32(projmatrix, "projmatrix")
        P = means3D.shape[0]
        present = torch.empty((P,), dtype=torch.bool, device=dev)
        if P:
            _check(L.goi_mark_visible(P, _ptr(m), _ptr(v), _ptr(p), present.data_ptr(), _stream(dev)),
                   "goi_mark_visible")
    return present


def hyperplane_mask(x, mlp_weight, mlp_bias, lut, hyperplane_w, hyperplane_b=0.0, log_scale=0.0, thresh=0.86,
                    mode=GOI_MASK_APE, channels_first=False, want_idx=False):
    """Fused GUI.compute_similarity (reference gui/main.py:363-385).

    x: [N,S] (channels_first=False) or planar [S,...] (channels_first=True, e.g. the render's [S,H,W]).
    Returns (sim[N], bg_mask[N] bool, idx[N] int32 or None)."""
    L = lib()
    dev = x.device
    with torch.cuda.device(dev):
        x = _f32(x, "x")
        if channels_first:
            S = x.shape[0]
            N = x.numel() // S if S else 0
            stride_n, stride_c = 1, N
        else:
            S = x.shape[-1]
            N = x.numel() // S if S else 0
            stride_n, stride_c = S, 1
        w, lut = _f32(mlp_weight, "mlp_weight"), _f32(lut, "lut")
        b = _f32(mlp_bias, "mlp_bias")
        hw = _f32(hyperplane_w.reshape(-1), "hyperplane_w")
        K, D = lut.shape
        if w.shape != (K, S):
            raise RuntimeError(f"mlp_weight must be [{K},{S}], got {tuple(w.shape)}")
        sim = torch.empty((N,), dtype=torch.float32, device=dev)
        bg = torch.empty((N,), dtype=torch.bool, device=dev)
        idx = torch.empty((N,), dtype=torch.int32, device=dev) if want_idx else None
        table = torch.empty((K,), dtype=torch.float32, device=dev)
        a = goi_mask_args(N, S, K, D, int(mode), stride_n, stride_c, _ptr(x), _ptr(w), _ptr(b), _ptr(lut), _ptr(hw),
                          float(hyperplane_b), float(log_scale), float(thresh), _ptr(table), _ptr(sim),
                          bg.data_ptr() if N else None, idx.data_ptr() if (want_idx and N) else None)
        _check(L.goi_mask(C.byref(a), _stream(dev)), "goi_mask")
    return sim, bg, idx


def rasterize_gaussians_mask(bg, means3D, colors, semantics, opacity, scales, rotations, scale_modifier,
                             cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh,
                             degree, campos, prefiltered, debug, mlp_weight, mlp_bias, lut, hyperplane_w,
                             hyperplane_b=0.0, log_scale=0.0, thresh=0.86, mode=GOI_MASK_APE, want_semantics=True,
                             want_idx=False):
    """Forward render + open-vocabulary mask in ONE pass (goi_forward_mask).  Inference
    only.  Returns (color, semantic or None, depth, alpha, radii, sim[H,W], bg_mask[H,W] bool, idx[H,W] or None);
    with want_semantics=False the [S,H,W] semantic image is never materialised."""
    L = lib()
    keep = []
    dev = means3D.device
    with torch.cuda.device(dev):
        g, means3D = _make_gaussians(means3D, sh, colors, semantics, opacity, scales, rotations, cov3D_precomp, keep)
        view = _make_view(bg, viewmatrix, projmatrix, campos, image_width, image_height, tan_fovx, tan_fovy,
                          scale_modifier, degree, prefiltered, debug, keep)
        if g.P == 0:                       # an empty [0,S] tensor travels as NULL: take S from the projection
            g.S = int(mlp_weight.shape[1])
        P, S, H, W = g.P, g.S, int(image_height), int(image_width)
        if S <= 0:
            raise RuntimeError("the fused mask needs semantic channels")
        f32 = dict(dtype=torch.float32, device=dev)
        out_color = torch.empty((3, H, W), **f32)
        out_sem = torch.empty((S, H, W), **f32) if want_semantics else None
        out_depth = torch.empty((1, H, W), **f32)
        out_alpha = torch.empty((1, H, W), **f32)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        w, lut_ = _f32(mlp_weight, "mlp_weight"), _f32(lut, "lut")
        b = _f32(mlp_bias, "mlp_bias")
        hw = _f32(hyperplane_w.reshape(-1), "hyperplane_w")
        K, D = lut_.shape
        if w.shape != (K, S):
            raise RuntimeError(f"mlp_weight must be [{K},{S}], got {tuple(w.shape)}")
        sim = torch.empty((H, W), **f32)
        bgm = torch.empty((H, W), dtype=torch.bool, device=dev)
        idx = torch.empty((H, W), dtype=torch.int32, device=dev) if want_idx else None
        table = torch.empty((K,), **f32)
        m = goi_mask_args(H * W, S, K, D, int(mode), 1, H * W, None, _ptr(w), _ptr(b), _ptr(lut_), _ptr(hw),
                          float(hyperplane_b), float(log_scale), float(thresh), _ptr(table), _ptr(sim),
                          bgm.data_ptr(), idx.data_ptr() if want_idx else None)
        u8 = dict(dtype=torch.uint8, device=dev)
        geom_bytes, img_bytes = L.goi_geom_bytes(P, S), L.goi_image_bytes(W, H)
        geom, img = torch.empty((geom_bytes,), **u8), torch.empty((img_bytes,), **u8)
        stream = _stream(dev)
        R = C.c_int64(0)
        out = goi_fwd_out(_ptr(out_color), _ptr(out_sem), _ptr(out_depth), _ptr(out_alpha), _ptr(radii))
        st = device_state(dev)
        bin_bytes = L.goi_binning_bytes(max(int(st.r_guess * 1.25), 4 * P, 1 << 16)) if P else L.goi_binning_bytes(0)
        binning = torch.empty((bin_bytes,), **u8)
        args = (C.byref(view), C.byref(g), C.byref(out), C.byref(m), geom.data_ptr(), geom_bytes)
        rc = L.goi_forward_mask(*args, binning.data_ptr(), bin_bytes, img.data_ptr(), img_bytes, stream, C.byref(R))
        if rc == -3 and R.value > 0 and L.goi_binning_bytes(R.value) > bin_bytes:
            bin_bytes = L.goi_binning_bytes(R.value)
            binning = torch.empty((bin_bytes,), **u8)
            rc = L.goi_forward_mask(*args, binning.data_ptr(), bin_bytes, img.data_ptr(), img_bytes, stream,
                                    C.byref(R))
        _check(rc, "goi_forward_mask")
        st.r_guess = max(R.value, int(0.9 * st.r_guess))
        st.last_num_rendered = int(R.value)
    return out_color, out_sem, out_depth, out_alpha, radii, sim, bgm, idx


def timing_enable(on: bool) -> None:
    _check(lib().goi_timing_enable(int(on)), "goi_timing_enable")


def timing_read() -> dict:
    """Mean ms per view of every pipeline stage over the (up to 64) views recorded since timing_enable(True)."""
    buf = (C.c_float * GOI_NUM_STAGES)()
    _check(lib().goi_timing_read(buf), "goi_timing_read")
    return {lib().goi_stage_name(i).decode(): float(buf[i]) for i in range(GOI_NUM_STAGES)}


def launch_count() -> int:
    return int(lib().goi_launch_count())


def read_stats(raster_settings_view, gaussians_struct, geom, radii, device):
    st = goi_stats()
    _check(lib().goi_read_stats(C.byref(raster_settings_view), C.byref(gaussians_struct), geom.data_ptr(),
                                radii.data_ptr(), _stream(device), C.byref(st)), "goi_read_stats")
    return st
