"""Drop-in replacement for the reference's ``diff_gaussian_rasterization`` Python package
(/root/reference/submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py).

Same public names, argument order, return tuples and error behaviour:

* ``GaussianRasterizationSettings``            reference :246-258
* ``GaussianRasterizer(nn.Module)`` with ``forward`` / ``markVisible`` / ``trace``   :260-349
* ``rasterize_gaussians`` / ``trace_gaussians``  :21-69
* ``_RasterizeGaussians(autograd.Function)``     :71-244 (saves the same 12 tensors, returns the
  gradients in the same order, keeps the debug snapshot dumps)

so ``gaussian_renderer.render`` (reference gaussian_renderer/__init__.py:14,86-95), ``train.py``,
``render.py`` and ``gui/gs_renderer.py`` import and call it unchanged.  Differences, all widening:
the semantic channel count is ``semantics.shape[1]`` at run time (the reference hard-codes
SEM_CHANNELS=10, cuda_rasterizer/config.h:18); ``semantics=None`` means S=0; kernels run on
torch's current stream under a device guard.  All compute happens in libgoi_raster.so (``_C.py``).
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _C


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, semantics, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, semantics, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


def trace_gaussians(means3D, means2D, sh, colors_precomp, img_sem, opacities, scales, rotations, cov3Ds_precomp,
                    raster_settings):
    return _RasterizeGaussians.trace(means3D, means2D, sh, colors_precomp, img_sem, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, semantics, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        # Restructure arguments the way the C++/CUDA side expects them (reference :86-107)
        args = (raster_settings.bg, means3D, colors_precomp, semantics, opacities, scales, rotations,
                raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy,
                raster_settings.image_height, raster_settings.image_width, sh, raster_settings.sh_degree,
                raster_settings.campos, raster_settings.prefiltered, raster_settings.debug)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)  # copy them before they can be corrupted
            try:
                (num_rendered, color, semant, depth, alpha, radii, geomBuffer, binningBuffer,
                 imgBuffer) = _C.rasterize_gaussians(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            (num_rendered, color, semant, depth, alpha, radii, geomBuffer, binningBuffer,
             imgBuffer) = _C.rasterize_gaussians(*args)
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.opacity_shape = opacities.shape
        ctx.save_for_backward(colors_precomp, semantics, means3D, scales, rotations, cov3Ds_precomp, radii, sh,
                              geomBuffer, binningBuffer, imgBuffer, alpha)
        ctx.mark_non_differentiable(radii)
        return color, semant, radii, depth, alpha

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_sem, grad_out_radii, grad_depth, grad_alpha):
        num_rendered = ctx.num_rendered
        raster_settings = ctx.raster_settings
        (colors_precomp, semantics, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
         binningBuffer, imgBuffer, alpha) = ctx.saved_tensors
        args = (raster_settings.bg, means3D, radii, colors_precomp, semantics, scales, rotations,
                raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy, grad_out_color,
                grad_out_sem, grad_depth, grad_alpha, sh, raster_settings.sh_degree, raster_settings.campos,
                geomBuffer, num_rendered, binningBuffer, imgBuffer, alpha, raster_settings.debug)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                (grad_means2D, grad_colors_precomp, grad_semantics, grad_opacities, grad_means3D,
                 grad_cov3Ds_precomp, grad_sh, grad_scales, grad_rotations) = _C.rasterize_gaussians_backward(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            (grad_means2D, grad_colors_precomp, grad_semantics, grad_opacities, grad_means3D, grad_cov3Ds_precomp,
             grad_sh, grad_scales, grad_rotations) = _C.rasterize_gaussians_backward(*args)

        def _shaped(grad, like):
            # absent optional inputs were `torch.Tensor([])`; autograd wants None (or a matching shape) for them
            if like is None or like.numel() == 0:
                return None
            return grad.view(like.shape) if grad.numel() == like.numel() else grad

        grads = (grad_means3D, grad_means2D, _shaped(grad_sh, sh), _shaped(grad_colors_precomp, colors_precomp),
                 _shaped(grad_semantics, semantics), grad_opacities.view(ctx.opacity_shape), _shaped(grad_scales, scales),
                 _shaped(grad_rotations, rotations), _shaped(grad_cov3Ds_precomp, cov3Ds_precomp), None)
        return grads

    @staticmethod
    def trace(means3D, means2D, sh, colors_precomp, img_sem, opacities, scales, rotations, cov3Ds_precomp,
              raster_settings):
        args = (raster_settings.bg, means3D, colors_precomp, img_sem, opacities, scales, rotations,
                raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy,
                raster_settings.image_height, raster_settings.image_width, sh, raster_settings.sh_degree,
                raster_settings.campos, raster_settings.prefiltered, raster_settings.debug)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                (num_rendered, color, gau_sem, num_gsem, geomBuffer, binningBuffer,
                 imgBuffer) = _C.rasterize_gaussians_trace(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            (num_rendered, color, gau_sem, num_gsem, geomBuffer, binningBuffer,
             imgBuffer) = _C.rasterize_gaussians_trace(*args)
        return color, gau_sem, num_gsem


class _RasterizeGaussiansRaw(torch.autograd.Function):
    """Not in the reference: the same rasterization taking the STORED parameters of
    scene/gaussian_model.py -- opacity logits, log-scales, un-normalised quaternions, SH split into
    _features_dc / _features_rest -- with sigmoid / exp / normalize / cat (get_opacity, get_scaling,
    get_rotation, get_features, :90-117) and their derivatives fused into the per-Gaussian kernels."""

    RAW = _C.GOI_RAW_OPACITY | _C.GOI_RAW_SCALE | _C.GOI_RAW_ROTATION

    @staticmethod
    def forward(ctx, means3D, means2D, features_dc, features_rest, semantics, opacity_logits, log_scales,
                raw_rotations, raster_settings):
        empty = torch.Tensor([])
        rs = raster_settings
        (num_rendered, color, semant, depth, alpha, radii, geomBuffer, binningBuffer, imgBuffer) = \
            _C.rasterize_gaussians(rs.bg, means3D, empty, semantics, opacity_logits, log_scales, raw_rotations,
                                   rs.scale_modifier, empty, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
                                   rs.image_height, rs.image_width, features_dc, rs.sh_degree, rs.campos,
                                   rs.prefiltered, rs.debug, raw_flags=_RasterizeGaussiansRaw.RAW,
                                   sh_rest=features_rest)
        ctx.raster_settings, ctx.num_rendered, ctx.opacity_shape = rs, num_rendered, opacity_logits.shape
        ctx.save_for_backward(semantics, means3D, log_scales, raw_rotations, radii, features_dc, features_rest,
                              geomBuffer, binningBuffer, imgBuffer, alpha)
        ctx.mark_non_differentiable(radii)
        return color, semant, radii, depth, alpha

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_sem, grad_out_radii, grad_depth, grad_alpha):
        rs = ctx.raster_settings
        (semantics, means3D, log_scales, raw_rotations, radii, features_dc, features_rest, geomBuffer, binningBuffer,
         imgBuffer, alpha) = ctx.saved_tensors
        empty = torch.Tensor([])
        (grad_means2D, _gc, grad_semantics, grad_opacities, grad_means3D, _gcov, grad_sh, grad_scales,
         grad_rotations) = _C.rasterize_gaussians_backward(
            rs.bg, means3D, radii, empty, semantics, log_scales, raw_rotations, rs.scale_modifier, empty,
            rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color, grad_out_sem, grad_depth,
            grad_alpha, features_dc, rs.sh_degree, rs.campos, geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer,
            alpha, rs.debug, raw_flags=_RasterizeGaussiansRaw.RAW, sh_rest=features_rest)
        if isinstance(grad_sh, tuple):
            grad_dc, grad_rest = grad_sh
        else:                       # degree-0 model: features_rest is [P,0,3] and travelled as NULL
            grad_dc, grad_rest = grad_sh, torch.zeros_like(features_rest)
        return (grad_means3D, grad_means2D, grad_dc.view(features_dc.shape), grad_rest.view(features_rest.shape),
                grad_semantics if semantics.numel() else None, grad_opacities.view(ctx.opacity_shape),
                grad_scales, grad_rotations, None)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _C.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    @staticmethod
    def _check_inputs(shs, colors_precomp, scales, rotations, cov3D_precomp):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, semantics=None, scales=None,
                rotations=None, cov3D_precomp=None):
        raster_settings = self.raster_settings
        self._check_inputs(shs, colors_precomp, scales, rotations, cov3D_precomp)
        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if semantics is None:
            semantics = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, semantics, opacities, scales, rotations,
                                   cov3D_precomp, raster_settings)

    def forward_raw(self, means3D, means2D, opacity_logits, features_dc, features_rest, log_scales, raw_rotations,
                    semantics=None):
        """Same outputs as ``forward`` from the STORED parameters (activations fused into the kernels):
        equals forward(means3D, means2D, sigmoid(opacity_logits), shs=cat((features_dc, features_rest), 1),
        semantics=semantics, scales=exp(log_scales), rotations=normalize(raw_rotations))."""
        if semantics is None:
            semantics = torch.Tensor([])
        return _RasterizeGaussiansRaw.apply(means3D, means2D, features_dc, features_rest, semantics, opacity_logits,
                                            log_scales, raw_rotations, self.raster_settings)

    @torch.no_grad()
    def forward_mask(self, means3D, opacities, hyperplane, shs=None, colors_precomp=None, semantics=None, scales=None,
                     rotations=None, cov3D_precomp=None, want_semantics=True, want_idx=False):
        """Inference render + open-vocabulary mask in one pass (not in the reference):
        equals ``forward(...)`` followed by ``hyperplane.compute_similarity(semantic_image)`` (gui/main.py:588-590,
        363-385) with the mask evaluated in the composite kernel's epilogue.  ``hyperplane`` is a
        ``goi_b200.semantic_mask.SemanticHyperplane``.  Returns a dict: render, semantics (None when
        want_semantics=False -- the [S,H,W] image is then never written), depth, alpha, radii, sim [H,W] (zeros
        below the threshold), bg_mask [H,W] bool, idx [H,W] int32 or None."""
        rs = self.raster_settings
        self._check_inputs(shs, colors_precomp, scales, rotations, cov3D_precomp)
        e = torch.Tensor([])
        mode, thresh, bias = hyperplane.kernel_args()
        color, sem, depth, alpha, radii, sim, bg, idx = _C.rasterize_gaussians_mask(
            rs.bg, means3D, e if colors_precomp is None else colors_precomp, semantics, opacities,
            e if scales is None else scales, e if rotations is None else rotations, rs.scale_modifier,
            e if cov3D_precomp is None else cov3D_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
            rs.image_height, rs.image_width, e if shs is None else shs, rs.sh_degree, rs.campos, rs.prefiltered,
            rs.debug, hyperplane.mlp_weight, hyperplane.mlp_bias, hyperplane.lut, hyperplane.w, hyperplane_b=bias,
            log_scale=hyperplane.log_scale, thresh=thresh, mode=mode, want_semantics=want_semantics,
            want_idx=want_idx)
        return {"render": color, "semantics": sem, "depth": depth, "alpha": alpha, "radii": radii, "sim": sim,
                "bg_mask": bg, "idx": idx}

    def trace(self, means3D, means2D, opacities, shs=None, colors_precomp=None, img_sem=None, scales=None,
              rotations=None, cov3D_precomp=None):
        raster_settings = self.raster_settings
        self._check_inputs(shs, colors_precomp, scales, rotations, cov3D_precomp)
        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if img_sem is None:
            img_sem = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        return trace_gaussians(means3D, means2D, shs, colors_precomp, img_sem, opacities, scales, rotations,
                               cov3D_precomp, raster_settings)
