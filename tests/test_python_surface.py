"""CPU tests of the host-side mirror of the reference's Python interface (names, argument checks,
error messages) and of the mask oracle against the reference's torch op chain."""
import math

import numpy as np
import pytest
import torch

from goi_b200.scenes import make_mask_model, make_scene
from common import torch_reference_similarity
from oracle import oracle


def test_settings_namedtuple_matches_reference_fields():
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    # reference: diff_gaussian_rasterization/__init__.py:246-258
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")


def test_module_surface():
    import diff_gaussian_rasterization as d
    for name in ("GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "trace_gaussians",
                 "_RasterizeGaussians", "cpu_deep_copy_tuple"):
        assert hasattr(d, name)
    for name in ("forward", "trace", "markVisible"):
        assert hasattr(d.GaussianRasterizer, name)
    for name in ("rasterize_gaussians", "rasterize_gaussians_backward", "rasterize_gaussians_trace", "mark_visible"):
        assert hasattr(d._C, name)          # ext.cpp:15-20
    import gaussian_renderer
    assert callable(gaussian_renderer.render) and callable(gaussian_renderer.trace)


def test_exactly_one_of_checks_raise_like_the_reference():
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    s = GaussianRasterizationSettings(8, 8, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 3,
                                      torch.zeros(3), False, False)
    r = GaussianRasterizer(s)
    z = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=z, means2D=z, opacities=torch.zeros(4, 1), scales=z, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=z, means2D=z, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 16, 3), colors_precomp=z, scales=z,
          rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(means3D=z, means2D=z, opacities=torch.zeros(4, 1), colors_precomp=z, scales=z)
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(means3D=z, means2D=z, opacities=torch.zeros(4, 1), colors_precomp=z, scales=z, rotations=torch.zeros(4, 4),
          cov3D_precomp=torch.zeros(4, 6))
    with pytest.raises(Exception, match="excatly one of either SHs"):
        r.trace(means3D=z, means2D=z, opacities=torch.zeros(4, 1), scales=z, rotations=torch.zeros(4, 4))


def test_cpu_tensors_are_rejected_not_silently_computed():
    """There is no CPU path: CPU inputs must raise, not fall back."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    s = GaussianRasterizationSettings(8, 8, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 3,
                                      torch.zeros(3), False, False)
    z = torch.zeros(4, 3)
    with pytest.raises(Exception):
        GaussianRasterizer(s)(means3D=z, means2D=z, opacities=torch.zeros(4, 1), colors_precomp=z, scales=z,
                              rotations=torch.zeros(4, 4))


@pytest.mark.parametrize("S,mode", [(16, "ape"), (10, "ape"), (32, "osh")])
def test_mask_oracle_matches_reference_torch_chain(S, mode):
    N = 5000
    gen = torch.Generator().manual_seed(S)
    x = torch.randn(N, S, generator=gen)
    mlp_w, mlp_b, lut, w = make_mask_model(S, seed=S)
    if mode == "ape":
        sim, bg, idx = torch_reference_similarity(x, mlp_w, mlp_b, lut, w, log_scale=0.3, thresh=0.86)
        o = oracle.mask(x.numpy(), mlp_w.numpy(), mlp_b.numpy(), lut.numpy(), w.numpy(), mode=0, log_scale=0.3, thresh=0.86)
    else:
        bias = 2 - math.log(0.86 / (1 - 0.86))
        sim, bg, idx = torch_reference_similarity(x, mlp_w, mlp_b, lut, w, osh_bias=bias)
        o = oracle.mask(x.numpy(), mlp_w.numpy(), mlp_b.numpy(), lut.numpy(), w.numpy(), mode=1, hyperplane_b=bias, thresh=0.5)
    clear = o["top2_gap"] > 1e-5                       # argmax near-ties are the documented exception
    assert clear.mean() > 0.99
    assert np.array_equal(o["idx"][clear], idx.numpy()[clear])
    assert np.abs(o["sim"][clear] - sim.numpy()[clear]).max() < 1e-5
    assert np.array_equal(o["bg_mask"][clear], bg.numpy()[clear])
    assert 0.02 < (~o["bg_mask"]).mean() < 0.98        # the random hyperplane actually splits the codebook
