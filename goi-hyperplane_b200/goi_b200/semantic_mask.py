"""Host-side mirror of the reference's mask path, backed by the fused CUDA kernel (goi_mask).

Reference: ``GUI.compute_similarity`` (gui/main.py:363-385) applied to the rendered semantic image
(``set_clip_mask`` :388-398) or to the per-Gaussian vectors (``compute_relative_gs_index`` :400-405).
The reference reshapes the render's planar [S,H,W] to [HW,S] first (gui/main.py:588); here the planar
layout is read in place (``channels_first=True``).
"""
from __future__ import annotations

import math

import torch

from diff_gaussian_rasterization import _C


def inverse_sigmoid(x: float) -> float:          # networks.py:10-11
    return math.log(x / (1 - x))


class SemanticHyperplane:
    """Bundles the learned pieces of the query: 1-layer MLP (S->K), codebook LUT [K,D], and either the
    APE text hyperplane (w, log_scale; ext/vision_language_align.py:109-122) or the fine-tuned
    LinearSVM (w, b; networks.py:12-59, 'OSH')."""

    def __init__(self, mlp_weight, mlp_bias, lut, text_feature, log_scale: float = 0.0, thresh: float = 0.86):
        self.mlp_weight, self.mlp_bias, self.lut = mlp_weight, mlp_bias, lut
        self.w = text_feature.reshape(-1)
        self.log_scale, self.thresh = float(log_scale), float(thresh)
        self.res_finetuned = False
        self.svm_bias = 0.0

    def enable_osh(self, weight=None, bias: float | None = None, set_bias: float = 0.86):
        """Switch to the LinearSVM head; default init = text vector and 2 - logit(0.86) (networks.py:18,
        gui/main.py:1677-1680)."""
        self.res_finetuned = True
        if weight is not None:
            self.w = weight.reshape(-1)
        self.svm_bias = (2 - inverse_sigmoid(set_bias)) if bias is None else float(bias)
        return self

    def kernel_args(self):
        """(mode, threshold, bias) of the active head: OSH uses 0.5 (gui/main.py:377-378), APE self.thresh."""
        if self.res_finetuned:
            return _C.GOI_MASK_OSH, 0.5, self.svm_bias
        return _C.GOI_MASK_APE, self.thresh, 0.0

    @torch.no_grad()
    def compute_similarity(self, embedding_feature, out_bg_mask=None, channels_first=False, want_idx=False):
        """gui/main.py:363-385: returns sim with below-threshold entries zeroed; fills out_bg_mask."""
        mode, thresh, bias = self.kernel_args()
        sim, bg, idx = _C.hyperplane_mask(embedding_feature, self.mlp_weight, self.mlp_bias, self.lut, self.w,
                                          hyperplane_b=bias, log_scale=self.log_scale, thresh=thresh, mode=mode,
                                          channels_first=channels_first, want_idx=want_idx)
        if out_bg_mask is not None:
            out_bg_mask[:] = bg
        return (sim, idx) if want_idx else sim

    @torch.no_grad()
    def mask_from_render(self, rendered_semantics):
        """[S,H,W] render output -> bool [H,W] mask (gui/main.py:396: mask = cos_sim > 0)."""
        S, H, W = rendered_semantics.shape
        sim = self.compute_similarity(rendered_semantics, channels_first=True)
        return (sim > 0).view(H, W)

    @torch.no_grad()
    def select_gaussians(self, gaussian_semantics):
        """[P,S] -> bool [P] (gui/main.py:400-405 compute_relative_gs_index)."""
        return self.compute_similarity(gaussian_semantics) > 0
