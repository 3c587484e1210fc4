"""Training-side semantic loss of GOI (reference train.py:142-170) as ONE fused CUDA call, forward + backward.

    loss, terms = semantic_loss(render_pkg["semantics"], semantic_MLP, lut, viewpoint_cam.semantic["ape"], iteration)
    loss.backward()        # gradients reach the rendered features (-> the rasterizer), the MLP and the codebook

replaces the reference's ~25 torch ops and their autograd graph (about ten [HW,300] temporaries + two [HW,256] ones per
iteration) with include/goi_semloss.h: both contractions are hand-written tcgen05 kernels (the similarity matrix lives in
tensor memory and never reaches HBM; the only [HW,300] array is its gradient) around two fused row kernels.  `sem_feature` is the render's planar [S,H,W] output and
`gt` the dataset's planar [D,H,W] target (both read in place; the reference permutes + reshapes both, train.py:142,147),
or [N,S] / [N,D] matrices.  There is no CPU/eager fallback: a missing library is an ImportError.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GOI_SEMLOSS_LIB", os.path.join(_HERE, "..", "lib", "libgoi_semloss.so"))
GOI_SEMLOSS_ABI_VERSION = 1
GOI_SEMLOSS_FP32, GOI_SEMLOSS_TF32 = 0, 1


class goi_semloss_args(C.Structure):
    _fields_ = [("N", C.c_int64), ("S", C.c_int32), ("K", C.c_int32), ("D", C.c_int32), ("precision", C.c_int32),
                ("anneal_t", C.c_float), ("_pad", C.c_int32), ("x", C.c_void_p), ("x_stride_n", C.c_int64),
                ("x_stride_c", C.c_int64), ("gt", C.c_void_p), ("gt_planar", C.c_int32), ("_pad2", C.c_int32),
                ("mlp_weight", C.c_void_p), ("mlp_bias", C.c_void_p), ("lut", C.c_void_p), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t), ("losses", C.c_void_p), ("dL_dx", C.c_void_p),
                ("dL_dmlp_weight", C.c_void_p), ("dL_dmlp_bias", C.c_void_p), ("dL_dlut", C.c_void_p)]


SYMBOLS = {
    "goi_semloss_abi_version": (C.c_int, []),
    "goi_semloss_last_error": (C.c_char_p, []),
    "goi_semloss_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "goi_semantic_loss": (C.c_int, [C.POINTER(goi_semloss_args), C.c_void_p]),
}
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.abspath(LIB_PATH)
        if not os.path.exists(path):
            raise ImportError(f"{path} not found: build it with `python goi-hyperplane_b200/build.py` "
                              "(nvcc sm_100a). The semantic loss has no CPU/eager fallback.")
        h = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        if h.goi_semloss_abi_version() != GOI_SEMLOSS_ABI_VERSION:
            raise ImportError(f"{path}: ABI {h.goi_semloss_abi_version()} != binding {GOI_SEMLOSS_ABI_VERSION}")
        _lib = h
    return _lib


def _layout(t, width_first_planar):
    """(N, width, stride_n, stride_c, planar) of a feature tensor: planar [C,H,W] / [C,N] or row-major [N,C]."""
    if width_first_planar:
        Cc = t.shape[0]
        N = t.numel() // Cc
        return N, Cc, 1, N, True
    Cc = t.shape[-1]
    return t.numel() // Cc, Cc, Cc, 1, False


class _SemanticLoss(torch.autograd.Function):
    """forward computes the loss AND its gradients in the same fused pass (the backward of this loss needs every
    intermediate of the forward; recomputing them later would double the work); backward scales by grad_output."""

    @staticmethod
    def forward(ctx, sem_feature, mlp_weight, mlp_bias, lut, gt, t, sem_planar, gt_planar, precision):
        L = lib()
        for name, v in (("sem_feature", sem_feature), ("mlp_weight", mlp_weight), ("lut", lut), ("gt", gt)):
            if not v.is_cuda or v.dtype != torch.float32:
                raise RuntimeError(f"{name} must be a float32 CUDA tensor")
        dev = sem_feature.device
        x, gtc = sem_feature.contiguous(), gt.contiguous()
        W, lutc = mlp_weight.contiguous(), lut.contiguous()
        b = None if mlp_bias is None else mlp_bias.contiguous()
        N, S, xs_n, xs_c, _ = _layout(x, sem_planar)
        Ng, D, _, _, _ = _layout(gtc, gt_planar)
        K = lutc.shape[0]
        if Ng != N:
            raise RuntimeError(f"sem_feature has {N} pixels, gt has {Ng}")
        if W.shape != (K, S) or lutc.shape[1] != D:
            raise RuntimeError(f"shape mismatch: mlp_weight {tuple(W.shape)}, lut {tuple(lutc.shape)}, S={S}, D={D}")
        with torch.cuda.device(dev):
            ws_bytes = L.goi_semloss_workspace_bytes(N, K, D)
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            losses = torch.empty((8,), dtype=torch.float32, device=dev)
            dx = torch.empty_like(x)
            dW, dlut = torch.empty_like(W), torch.empty_like(lutc)
            db = None if b is None else torch.empty_like(b)
            a = goi_semloss_args(N, S, K, D, int(precision), float(t), 0, x.data_ptr(), xs_n, xs_c, gtc.data_ptr(),
                                 int(gt_planar), 0, W.data_ptr(), None if b is None else b.data_ptr(),
                                 lutc.data_ptr(), ws.data_ptr(), ws_bytes, losses.data_ptr(), dx.data_ptr(),
                                 dW.data_ptr(), None if db is None else db.data_ptr(), dlut.data_ptr())
            rc = L.goi_semantic_loss(C.byref(a), torch.cuda.current_stream(dev).cuda_stream)
            if rc != 0:
                raise RuntimeError(f"goi_semantic_loss failed ({rc}): {L.goi_semloss_last_error().decode()}")
        ctx.save_for_backward(dx, dW, dlut, *([] if db is None else [db]))
        ctx.has_bias = db is not None
        ctx.mark_non_differentiable(losses)
        return losses[0].clone(), losses

    @staticmethod
    def backward(ctx, grad_loss, _grad_terms):
        saved = ctx.saved_tensors
        dx, dW, dlut = saved[0], saved[1], saved[2]
        db = saved[3] if ctx.has_bias else None
        g = grad_loss
        return (dx * g, dW * g, None if db is None else db * g, dlut * g, None, None, None, None, None)


def semantic_loss(sem_feature, semantic_mlp, lut, gt, iteration=1, sem_planar=None, gt_planar=None,
                  precision=GOI_SEMLOSS_FP32):
    """train.py:142-163.  semantic_mlp: the reference's SemanticModel(num_layer=1) / an nn.Linear / a (weight, bias)
    pair.  sem_feature: [S,H,W] (planar, default for 3-D input) or [N,S]; gt: [D,H,W] or [N,D].
    Returns (loss, terms) with terms = [loss, lab, sl, sl1, recc, min(sim_val), 0, 0] (device tensor, no grad)."""
    if isinstance(semantic_mlp, (tuple, list)):
        weight, bias = semantic_mlp
    else:
        lin = semantic_mlp.layers[0] if hasattr(semantic_mlp, "layers") else semantic_mlp
        weight, bias = lin.weight, lin.bias
    if sem_planar is None:
        sem_planar = sem_feature.ndim == 3
    if gt_planar is None:
        gt_planar = gt.ndim == 3
    t = 1.0 if iteration < 1000 else 2.0                      # train.py:157
    return _SemanticLoss.apply(sem_feature, weight, bias, lut, gt, t, bool(sem_planar), bool(gt_planar), int(precision))
