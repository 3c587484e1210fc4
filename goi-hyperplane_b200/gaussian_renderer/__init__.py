"""Mirror of the reference's render wrapper (/root/reference/gaussian_renderer/__init__.py:18-192):
``render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None)`` and
``trace(viewpoint_camera, pc, img_sem, pipe, bg_color, ...)`` with the same marshalling, the same
result dictionaries and the same two optional Python routes (``pipe.compute_cov3D_python``,
``pipe.convert_SHs_python``).  The reference's own file works unchanged on top of this repo's
``diff_gaussian_rasterization``; this copy only drops its imports of ``scene.gaussian_model`` and
``utils.sh_utils`` so that tests / bench need nothing but a duck-typed Gaussian container
(``get_xyz, get_opacity, get_scaling, get_rotation, get_features, get_semantics,
active_sh_degree, max_sh_degree, get_covariance``).
"""
import math

import torch

from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer

_C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)


def eval_sh(deg, sh, dirs):
    """Real SH basis up to degree 3 (what utils/sh_utils.py:57-112 evaluates): sh [...,C,(deg+1)^2], dirs [...,3]."""
    result = _C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = result - _C1 * y * sh[..., 1] + _C1 * z * sh[..., 2] - _C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            result = (result + _C2[0] * xy * sh[..., 4] + _C2[1] * yz * sh[..., 5]
                      + _C2[2] * (2.0 * zz - xx - yy) * sh[..., 6] + _C2[3] * xz * sh[..., 7]
                      + _C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                result = (result + _C3[0] * y * (3 * xx - yy) * sh[..., 9] + _C3[1] * xy * z * sh[..., 10]
                          + _C3[2] * y * (4 * zz - xx - yy) * sh[..., 11]
                          + _C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                          + _C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + _C3[5] * z * (xx - yy) * sh[..., 14]
                          + _C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return result


def _settings(viewpoint_camera, pc, pipe, bg_color, scaling_modifier):
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    return GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx,
        tanfovy=tanfovy,
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        debug=pipe.debug)


def _marshal(viewpoint_camera, pc, pipe, scaling_modifier, override_color):
    scales = rotations = cov3D_precomp = None
    if pipe.compute_cov3D_python:
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        scales = pc.get_scaling
        rotations = pc.get_rotation
    shs = colors_precomp = None
    if override_color is None:
        if pipe.convert_SHs_python:
            shs_view = pc.get_features.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
            dir_pp = (pc.get_xyz - viewpoint_camera.camera_center.repeat(pc.get_features.shape[0], 1))
            dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
            sh2rgb = eval_sh(pc.active_sh_degree, shs_view, dir_pp_normalized)
            colors_precomp = torch.clamp_min(sh2rgb + 0.5, 0.0)
        else:
            shs = pc.get_features
    else:
        colors_precomp = override_color
    return scales, rotations, cov3D_precomp, shs, colors_precomp


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None,
           fused_activations=False):
    """Render the scene (reference :18-105).  Background tensor (bg_color) must be on GPU!

    fused_activations=True (opt-in, not in the reference) hands the model's STORED
    parameters (``pc._opacity, _scaling, _rotation, _features_dc, _features_rest``) to the rasterizer, which
    applies sigmoid / exp / normalize / cat while loading them -- same images, gradients w.r.t. the stored
    parameters, none of the per-render activation kernels and no 384 B/Gaussian SH concatenation.  It needs the
    default pipeline (no override_color, no compute_cov3D_python / convert_SHs_python, no semantic masks)."""
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True,
                                          device=pc.get_xyz.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    rasterizer = GaussianRasterizer(raster_settings=_settings(viewpoint_camera, pc, pipe, bg_color, scaling_modifier))
    if fused_activations:
        if override_color is not None or pipe.compute_cov3D_python or pipe.convert_SHs_python:
            raise ValueError("fused_activations needs the default pipeline (SH colours, scale/rotation covariance)")
        if getattr(pc, "_semantics_masks", None) is not None:
            raise ValueError("fused_activations does not apply semantic masks; call set_semantic_masks(None)")
        rendered_image, rendered_sem, radii, depth, alpha = rasterizer.forward_raw(
            means3D=pc._xyz, means2D=screenspace_points, opacity_logits=pc._opacity, features_dc=pc._features_dc,
            features_rest=pc._features_rest, log_scales=pc._scaling, raw_rotations=pc._rotation,
            semantics=pc._semantics)
        return {"render": rendered_image, "semantics": rendered_sem, "depth": depth, "alpha": alpha,
                "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii}
    scales, rotations, cov3D_precomp, shs, colors_precomp = _marshal(viewpoint_camera, pc, pipe, scaling_modifier,
                                                                     override_color)
    rendered_image, rendered_sem, radii, depth, alpha = rasterizer(
        means3D=pc.get_xyz,
        means2D=screenspace_points,
        shs=shs,
        colors_precomp=colors_precomp,
        semantics=pc.get_semantics,
        opacities=pc.get_opacity,
        scales=scales,
        rotations=rotations,
        cov3D_precomp=cov3D_precomp)
    return {"render": rendered_image,
            "semantics": rendered_sem,
            "depth": depth,
            "alpha": alpha,
            "viewspace_points": screenspace_points,
            "visibility_filter": radii > 0,
            "radii": radii}


@torch.no_grad()
def render_mask(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, hyperplane, scaling_modifier=1.0,
                override_color=None, want_semantics=True, want_idx=False):
    """``render`` + ``GUI.compute_similarity`` on the rendered semantic image in one pass (the GUI's test_step,
    reference gui/main.py:575-600 with :363-385), using the composite kernel's fused mask epilogue.  Inference
    only.  Adds ``sim`` [H,W], ``bg_mask`` [H,W] and ``mask`` (= sim > 0, gui/main.py:396) to the render dict;
    ``semantics`` is None when want_semantics=False."""
    rasterizer = GaussianRasterizer(raster_settings=_settings(viewpoint_camera, pc, pipe, bg_color, scaling_modifier))
    scales, rotations, cov3D_precomp, shs, colors_precomp = _marshal(viewpoint_camera, pc, pipe, scaling_modifier,
                                                                     override_color)
    out = rasterizer.forward_mask(means3D=pc.get_xyz, opacities=pc.get_opacity, hyperplane=hyperplane, shs=shs,
                                  colors_precomp=colors_precomp, semantics=pc.get_semantics, scales=scales,
                                  rotations=rotations, cov3D_precomp=cov3D_precomp, want_semantics=want_semantics,
                                  want_idx=want_idx)
    out["visibility_filter"] = out["radii"] > 0
    out["mask"] = out["sim"] > 0
    return out


def trace(viewpoint_camera, pc, img_sem: torch.Tensor, pipe, bg_color: torch.Tensor, scaling_modifier=1.0,
          override_color=None):
    """Back-project a 2D feature image onto the Gaussians (reference :107-192)."""
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, device=pc.get_xyz.device)
    rasterizer = GaussianRasterizer(raster_settings=_settings(viewpoint_camera, pc, pipe, bg_color, scaling_modifier))
    scales, rotations, cov3D_precomp, shs, colors_precomp = _marshal(viewpoint_camera, pc, pipe, scaling_modifier,
                                                                     override_color)
    rendered_image, gau_sem, num_gsem = rasterizer.trace(
        means3D=pc.get_xyz,
        means2D=screenspace_points,
        shs=shs,
        colors_precomp=colors_precomp,
        img_sem=img_sem,
        opacities=pc.get_opacity,
        scales=scales,
        rotations=rotations,
        cov3D_precomp=cov3D_precomp)
    return {"render": rendered_image,
            "gaussian_semantics": gau_sem,
            "num_gsem": num_gsem}
