"""the reference's OWN callers, source unchanged, running on this package.

Loaded by path from /root/reference (authoring container only -- the GPU box has no reference tree, so these tests
skip there):
  * gaussian_renderer/__init__.py:18-105   render()            (train.py:132, render.py:21)
  * gui/gs_renderer.py:231-348             Renderer.render()   incl. the `gaussian_mask` index subset (gui/main.py:1878)
  * scene/gaussian_model.py:13-130         the real GaussianModel with its activation getters (sigmoid / exp /
                                           normalize / cat), utils/sh_utils.py eval_sh, utils/general_utils.py
Only modules OUTSIDE the path are stubbed (plyfile, simple_knn, mesh, mesh_utils, kiui, scene/__init__.py's dataset
readers).  `from diff_gaussian_rasterization import ...` inside those files resolves to THIS repo's package.

There is no GPU in the authoring container and the product has no CPU path, so on a CPU-only host the four `_C`
entry points are backed by the CPU oracle (test infrastructure; tests may use it) and the literal `device="cuda"` in
the reference files is redirected by a proxy of the `torch` name in those modules' globals.  Everything above `_C`
-- GaussianRasterizationSettings, GaussianRasterizer.forward's checks and None handling, _RasterizeGaussians'
argument marshalling, saved tensors and gradient order -- is this repo's product code, driven by the reference's
callers.  With a GPU and the reference tree present the same tests run on the real library.
"""
import contextlib
import importlib.util
import math
import os
import sys
import types

import numpy as np
import pytest
import torch

import common
from goi_b200.scenes import PipeFlags, SyntheticCamera, make_loss_weights, make_scene

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "gaussian_renderer")),
                                reason="reference tree not present (GPU box)")
ON_GPU = torch.cuda.is_available()
DEV = "cuda" if ON_GPU else "cpu"


class _TorchCudaToCpu:
    """Stands in for the name `torch` inside the reference modules on a CPU-only host: device="cuda" -> "cpu"."""

    def __init__(self):
        self._t = torch

    def __getattr__(self, name):
        attr = getattr(self._t, name)
        if name in ("zeros_like", "tensor", "zeros", "ones", "empty"):
            def wrapped(*a, **k):
                if str(k.get("device", "")) .startswith("cuda"):
                    k["device"] = "cpu"
                return attr(*a, **k)
            return wrapped
        return attr


def _load(name, path, stub_torch=True):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    if stub_torch and not ON_GPU:
        mod.torch = _TorchCudaToCpu()
    return mod


@contextlib.contextmanager
def reference_modules():
    """Import the reference's caller files unchanged; restore sys.modules / sys.path afterwards."""
    saved_modules, saved_path = dict(sys.modules), list(sys.path)
    try:
        def stub(name, **attrs):
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
            return m
        # out-of-scope dependencies of the files on the path
        stub("plyfile", PlyData=object, PlyElement=object)
        stub("simple_knn")
        stub("simple_knn._C", distCUDA2=None)
        stub("mesh", Mesh=object)
        stub("mesh_utils", decimate_mesh=None, clean_mesh=None)
        stub("kiui")
        # the reference's own pure-torch helpers, real: utils/{general_utils,system_utils,sh_utils,graphics_utils}.py
        utils = stub("utils")
        utils.__path__ = [os.path.join(REF, "utils")]
        # `scene` as a bare package so scene/__init__.py (dataset readers, argparse, PIL) does not run
        scene = stub("scene")
        scene.__path__ = [os.path.join(REF, "scene")]
        gm = importlib.import_module("scene.gaussian_model")
        sm = importlib.import_module("scene.semantic_model")
        scene.GaussianModel, scene.SemanticModel = gm.GaussianModel, sm.SemanticModel
        if not ON_GPU:                  # utils/general_utils.py:76,94,113 also spell device="cuda"
            for name, m in list(sys.modules.items()):
                if name.startswith(("utils.", "scene.")) and hasattr(m, "torch"):
                    m.torch = _TorchCudaToCpu()
        ref_render = _load("ref_gaussian_renderer", os.path.join(REF, "gaussian_renderer", "__init__.py"))
        ref_gui = _load("ref_gs_renderer", os.path.join(REF, "gui", "gs_renderer.py"))
        import diff_gaussian_rasterization as ours
        assert ref_render.GaussianRasterizer is ours.GaussianRasterizer        # the import resolved to this repo
        assert ref_gui.GaussianRasterizationSettings is ours.GaussianRasterizationSettings
        yield types.SimpleNamespace(render=ref_render.render, Renderer=ref_gui.Renderer, GaussianModel=gm.GaussianModel,
                                    gs_module=ref_gui)
    finally:
        for k in list(sys.modules):
            if k not in saved_modules:
                del sys.modules[k]
        sys.modules.update(saved_modules)
        sys.path[:] = saved_path


@contextlib.contextmanager
def c_backend(calls):
    """On a CPU-only host: _C.rasterize_gaussians / _backward served by the CPU oracle, recording their arguments."""
    if ON_GPU:
        yield
        return
    from diff_gaussian_rasterization import _C
    from oracle import oracle
    held = {}
    n = lambda t: None if t is None or t.numel() == 0 else t.detach().cpu().numpy()

    def fwd(bg, means3D, colors, semantics, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
            projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered, debug):
        calls.append(dict(fn="rasterize_gaussians", P=means3D.shape[0], S=semantics.shape[1] if semantics.numel() else 0,
                          sh=tuple(sh.shape), scales=tuple(scales.shape), cov=tuple(cov3D_precomp.shape),
                          colors=tuple(colors.shape), degree=degree, HW=(image_height, image_width)))
        res = oracle.forward(means3D=n(means3D), opacities=n(opacity), shs=n(sh), colors_precomp=n(colors),
                             semantics=n(semantics), scales=n(scales), rotations=n(rotations),
                             cov3D_precomp=n(cov3D_precomp), W=image_width, H=image_height, viewmatrix=n(viewmatrix),
                             projmatrix=n(projmatrix), campos=n(campos), tanfovx=tan_fovx, tanfovy=tan_fovy, bg=n(bg),
                             sh_degree=degree, scale_modifier=scale_modifier)
        key = torch.tensor([len(held)], dtype=torch.int64)
        held[len(held)] = res
        t = torch.from_numpy
        return (res.num_rendered, t(res.color), t(res.semantics), t(res.depth), t(res.alpha), t(res.radii),
                key, torch.zeros(1, dtype=torch.uint8), torch.zeros(1, dtype=torch.uint8))

    def bwd(bg, means3D, radii, colors, semantics, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
            projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_semantic, dL_dout_depth, dL_dout_alpha, sh, degree,
            campos, geomBuffer, R, binningBuffer, imageBuffer, alphas, debug):
        calls.append(dict(fn="rasterize_gaussians_backward", R=R))
        res = held[int(geomBuffer[0])]
        g = oracle.backward(res, n(dL_dout_color), n(dL_dout_semantic), n(dL_dout_depth), n(dL_dout_alpha), wide=True)
        t = torch.from_numpy
        return (t(g["dL_dmeans2D"]), t(g["dL_dcolors"]), t(g["dL_dsemantics"]), t(g["dL_dopacity"]), t(g["dL_dmeans3D"]),
                t(g["dL_dcov3D"]), t(g["dL_dsh"]), t(g["dL_dscales"]), t(g["dL_drotations"]))

    saved = _C.rasterize_gaussians, _C.rasterize_gaussians_backward
    _C.rasterize_gaussians, _C.rasterize_gaussians_backward = fwd, bwd
    try:
        yield
    finally:
        _C.rasterize_gaussians, _C.rasterize_gaussians_backward = saved


def _reference_model(ref, g, S, raw_gain=None):
    """The reference's real GaussianModel holding the STORED parameters whose activations reproduce scene `g`."""
    pc = ref.GaussianModel(3, S)
    pc.active_sh_degree = 3
    P = g.get_xyz.shape[0]
    gain = torch.ones(P, 1) if raw_gain is None else raw_gain
    leaf = lambda t: torch.nn.Parameter(t.clone().to(DEV).contiguous())
    pc._xyz = leaf(g.get_xyz)
    pc._features_dc = leaf(g.get_features[:, :1])
    pc._features_rest = leaf(g.get_features[:, 1:])
    pc._semantics = leaf(g.get_semantics)
    pc._scaling = leaf(torch.log(g.get_scaling))
    pc._rotation = leaf(g.get_rotation * gain)              # |q| != 1: get_rotation's normalize matters
    pc._opacity = leaf(torch.logit(g.get_opacity))
    return pc


def _activated(pc):
    """What the reference's getters hand to the rasterizer, as a SyntheticGaussians for the direct oracle run."""
    from goi_b200.scenes import SyntheticGaussians
    with torch.no_grad():
        return SyntheticGaussians(pc.get_xyz.cpu(), pc.get_opacity.cpu(), pc.get_scaling.cpu(), pc.get_rotation.cpu(),
                                  pc.get_features.cpu().contiguous(), pc.get_semantics.cpu())


@pytest.mark.parametrize("convert_SHs_python,compute_cov3D_python", [(False, False), (True, False), (False, True)])
def test_reference_render_runs_unchanged_on_this_package(convert_SHs_python, compute_cov3D_python):
    """gaussian_renderer/__init__.py:18-105 with the reference's own GaussianModel, forward + backward; including the
    two Python routes the reference offers as cross-checks of the device math (pipe.convert_SHs_python ->
    utils/sh_utils.eval_sh, pipe.compute_cov3D_python -> build_covariance_from_scaling_rotation)."""
    P, W, H, S = 1500, 96, 64, 10
    g, cam, _ = make_scene(P, W, H, S, 31)
    cam, bg = cam.to(DEV), torch.tensor([0.2, 0.4, 0.1], device=DEV)
    w = make_loss_weights(S, W, H, 31, device=DEV)
    calls = []
    with reference_modules() as ref, c_backend(calls):
        gain = 0.5 + 2.0 * torch.rand(P, 1, generator=torch.Generator().manual_seed(1))
        pc = _reference_model(ref, g, S, gain)
        pipe = PipeFlags(convert_SHs_python=convert_SHs_python, compute_cov3D_python=compute_cov3D_python)
        out = ref.render(cam, pc, pipe, bg)
        assert list(out) == ["render", "semantics", "depth", "alpha", "viewspace_points", "visibility_filter", "radii"]
        assert out["render"].shape == (3, H, W) and out["semantics"].shape == (S, H, W)
        assert out["depth"].shape == (1, H, W) and out["alpha"].shape == (1, H, W)
        assert out["radii"].shape == (P,) and out["radii"].dtype == torch.int32
        assert out["visibility_filter"].dtype == torch.bool and 0 < int(out["visibility_filter"].sum()) <= P
        loss = sum((out[k] * w[k]).sum() for k in ("render", "semantics", "depth", "alpha"))
        loss.backward()                                     # train.py:168
        grads = {k: getattr(pc, k).grad for k in ("_xyz", "_features_dc", "_features_rest", "_semantics", "_scaling",
                                                  "_rotation", "_opacity")}
        assert all(v is not None and torch.isfinite(v).all() for v in grads.values())
        assert out["viewspace_points"].grad is not None and out["viewspace_points"].grad.shape == (P, 3)
        assert float(out["viewspace_points"].grad[:, 2].abs().max()) == 0.0          # dmeans2D.z == 0 (PY:176-187)

        # the same scene straight through the oracle on the activated tensors (default route: SHs + scale/rotation)
        ga = _activated(pc)
        ora = common.run_oracle(ga, cam, bg, w)
        tol = 0 if not (convert_SHs_python or compute_cov3D_python or ON_GPU) else None
        got = {"color": out["render"], "semantics": out["semantics"], "depth": out["depth"], "alpha": out["alpha"]}
        if tol == 0:
            for k in got:
                assert np.array_equal(common.to_np(got[k]), ora[k]), k
            assert np.array_equal(common.to_np(out["radii"]), ora["radii"])
        else:
            # the Python routes round differently from the device math (radii may flip by one on a few splats)
            assert (common.to_np(out["radii"]) != ora["radii"]).mean() < 2e-3
            common.assert_images_close(got, ora, max_bad_frac=2e-3, what="reference python route vs device route")
        rel = float((out["viewspace_points"].grad.cpu() - torch.from_numpy(ora["grads"]["dL_dmeans2D"])).abs().max()) \
            / float(np.abs(ora["grads"]["dL_dmeans2D"]).max())
        assert rel <= (1e-6 if tol == 0 else 5e-3)
        # dL/d_semantics needs no activation chain: equal to the oracle's dL_dsemantics
        rel = float((grads["_semantics"].cpu() - torch.from_numpy(ora["grads"]["dL_dsemantics"])).abs().max()) \
            / float(np.abs(ora["grads"]["dL_dsemantics"]).max())
        assert rel <= (1e-6 if tol == 0 else 5e-3)

        # this repo's mirror of the wrapper must behave identically on the same model
        from gaussian_renderer import render as mirror_render
        for p in grads:
            getattr(pc, p).grad = None
        out2 = mirror_render(cam, pc, pipe, bg)
        assert list(out2) == list(out)
        for k in ("render", "semantics", "depth", "alpha", "radii", "visibility_filter"):
            assert torch.equal(out[k], out2[k]), k
    if not ON_GPU:
        f = [c for c in calls if c["fn"] == "rasterize_gaussians"]
        assert len(f) == 2 and f[0] == f[1]
        # absent optional inputs arrive as empty tensors (PY:286-297), exactly one of each pair is populated
        assert (f[0]["sh"] == (0,)) == convert_SHs_python and (f[0]["colors"] == (0,)) != convert_SHs_python
        assert (f[0]["scales"] == (0,)) == compute_cov3D_python and (f[0]["cov"] == (0,)) != compute_cov3D_python
        assert f[0]["S"] == S and f[0]["degree"] == 3 and f[0]["HW"] == (H, W)


def test_reference_gui_renderer_with_gaussian_mask():
    """gui/gs_renderer.py:231-348 Renderer.render(), no_grad, with and without the `gaussian_mask` subset the GUI uses
    for 3D retrieve / delete (gui/main.py:1878).  The reference indexes every per-Gaussian tensor EXCEPT means2D with
    the mask, so the binding must accept a screen-space tensor of a different length."""
    P, W, H, S = 1200, 80, 56, 10                          # gs_renderer.py:179 hard-codes 10 semantic channels
    g, cam, _ = make_scene(P, W, H, S, 37)
    cam = cam.to(DEV)
    calls = []
    with reference_modules() as ref, c_backend(calls):
        r = ref.Renderer(sh_degree=3, white_background=True)
        assert float(r.bg_color.sum()) == 3.0
        r.gaussians = _reference_model(ref, g, S)
        ga = _activated(r.gaussians)
        with torch.no_grad():
            full = r.render(cam)
            keep = torch.rand(P, generator=torch.Generator().manual_seed(3)) < 0.4
            sub = r.render(cam, gaussian_mask=keep.to(DEV), bg_color=torch.tensor([0.1, 0.2, 0.3], device=DEV))
        assert list(full) == ["image", "semantics", "depth", "alpha", "viewspace_points", "visibility_filter", "radii"]
        ora = common.run_oracle(ga, cam, torch.ones(3))
        assert float((full["image"].cpu() - torch.from_numpy(ora["color"]).clamp(0, 1)).abs().max()) <= (1e-4 if ON_GPU else 0)
        assert np.array_equal(common.to_np(full["radii"]), ora["radii"]) or ON_GPU
        from goi_b200.scenes import SyntheticGaussians
        gs = SyntheticGaussians(*[t[keep].contiguous() for t in ga.tensors()])
        ora_s = common.run_oracle(gs, cam, torch.tensor([0.1, 0.2, 0.3]))
        assert sub["radii"].shape == (int(keep.sum()),) and sub["viewspace_points"].shape == (P, 3)
        assert float((sub["image"].cpu() - torch.from_numpy(ora_s["color"]).clamp(0, 1)).abs().max()) <= (1e-4 if ON_GPU else 0)
        assert float((sub["semantics"].cpu() - torch.from_numpy(ora_s["semantics"])).abs().max()) <= (1e-4 if ON_GPU else 0)
    if not ON_GPU:
        f = [c for c in calls if c["fn"] == "rasterize_gaussians"]
        assert [c["P"] for c in f] == [P, int(keep.sum())]


def test_reference_caller_errors_surface_unchanged():
    """The exactly-one-of checks the reference's wrapper relies on (PY:280-284) raise the same messages."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    cam = SyntheticCamera(32, 32, math.radians(60))
    rs = GaussianRasterizationSettings(32, 32, 0.5, 0.5, torch.zeros(3), 1.0, cam.world_view_transform,
                                       cam.full_proj_transform, 3, cam.camera_center, False, False)
    z = torch.zeros
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        GaussianRasterizer(rs)(means3D=z(4, 3), means2D=z(4, 3), opacities=z(4, 1), scales=z(4, 3), rotations=z(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        GaussianRasterizer(rs)(means3D=z(4, 3), means2D=z(4, 3), opacities=z(4, 1), shs=z(4, 16, 3), scales=z(4, 3))
