"""Shared helpers of the parity tests: run one scene through (a) the CPU oracle, (b) this repo's CUDA
path via the drop-in Python surface, (c) the reference's own CUDA core (oracle/_ref), and compare."""
from __future__ import annotations

import math

import numpy as np
import torch

from goi_b200.scenes import PipeFlags, make_loss_weights, make_scene

IMG_TOL = 1e-4          # north_star: L-inf on rendered RGB / feature / mask buffers
GRAD_RTOL = 1e-3        # north_star: relative on gradients (of the tensor's max magnitude)

GRAD_KEYS = ("dL_dmeans3D", "dL_dmeans2D", "dL_dsh", "dL_dsemantics", "dL_dopacity", "dL_dscales", "dL_drotations")


def cam_arrays(cam, bg):
    return dict(W=cam.image_width, H=cam.image_height,
                viewmatrix=cam.world_view_transform.detach().cpu().numpy(),
                projmatrix=cam.full_proj_transform.detach().cpu().numpy(),
                campos=cam.camera_center.detach().cpu().numpy(),
                tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5),
                bg=bg.detach().cpu().numpy())


def gaussian_arrays(g, use_sh=True, use_cov=False):
    n = lambda t: None if t is None else t.detach().cpu().numpy()
    d = dict(means3D=n(g.get_xyz), opacities=n(g.get_opacity), semantics=n(g.get_semantics))
    if use_sh:
        d["shs"] = n(g.get_features)
    else:
        d["colors_precomp"] = n(torch.sigmoid(g.get_features[:, 0, :]))
    if use_cov:
        d["cov3D_precomp"] = n(g.get_covariance(1.0))
    else:
        d["scales"], d["rotations"] = n(g.get_scaling), n(g.get_rotation)
    return d


def run_oracle(g, cam, bg, weights=None, use_sh=True, use_cov=False, wide=True, sh_degree=3):
    from oracle import oracle
    res = oracle.forward(**gaussian_arrays(g, use_sh, use_cov), **cam_arrays(cam, bg), sh_degree=sh_degree)
    out = dict(color=res.color, semantics=res.semantics, depth=res.depth, alpha=res.alpha, radii=res.radii,
               num_rendered=res.num_rendered)
    if weights is not None:
        w = {k: v.detach().cpu().numpy() for k, v in weights.items()}
        out["grads"] = oracle.backward(res, w["render"], w["semantics"], w["depth"], w["alpha"], wide=wide)
    out["_res"] = res
    return out


def run_cuda(g, cam, bg, weights=None, use_sh=True, use_cov=False, device="cuda", sh_degree=3):
    """This repo's CUDA path through the reference-shaped Python API (GaussianRasterizer + autograd)."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    g = g.to(device)
    cam = cam.to(device)
    bg = bg.to(device)
    settings = GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=math.tan(cam.FoVx * 0.5),
        tanfovy=math.tan(cam.FoVy * 0.5), bg=bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform,
        projmatrix=cam.full_proj_transform, sh_degree=sh_degree, campos=cam.camera_center, prefiltered=False,
        debug=False)
    rast = GaussianRasterizer(settings)
    leaf = lambda t: None if t is None else t.detach().clone().requires_grad_(weights is not None)
    means3D, opacity, sem = leaf(g.get_xyz), leaf(g.get_opacity), leaf(g.get_semantics)
    means2D = torch.zeros_like(means3D, requires_grad=weights is not None)
    kw = {}
    if use_sh:
        kw["shs"] = leaf(g.get_features)
    else:
        kw["colors_precomp"] = leaf(torch.sigmoid(g.get_features[:, 0, :]))
    if use_cov:
        kw["cov3D_precomp"] = leaf(g.get_covariance(1.0))
    else:
        kw["scales"], kw["rotations"] = leaf(g.get_scaling), leaf(g.get_rotation)
    color, semantics, radii, depth, alpha = rast(means3D=means3D, means2D=means2D, opacities=opacity,
                                                 semantics=sem, **kw)
    out = dict(color=color, semantics=semantics, depth=depth, alpha=alpha, radii=radii)
    if weights is not None:
        loss = (color * weights["render"].to(device)).sum() + (depth * weights["depth"].to(device)).sum() \
            + (alpha * weights["alpha"].to(device)).sum()
        if semantics.numel():
            loss = loss + (semantics * weights["semantics"].to(device)).sum()
        loss.backward()
        gr = dict(dL_dmeans3D=means3D.grad, dL_dmeans2D=means2D.grad, dL_dopacity=opacity.grad,
                  dL_dsemantics=None if sem is None else sem.grad)
        if use_sh:
            gr["dL_dsh"] = kw["shs"].grad
        else:
            gr["dL_dcolors"] = kw["colors_precomp"].grad
        if use_cov:
            gr["dL_dcov3D"] = kw["cov3D_precomp"].grad
        else:
            gr["dL_dscales"], gr["dL_drotations"] = kw["scales"].grad, kw["rotations"].grad
        out["grads"] = gr
    return out


def run_reference_cuda(g, cam, bg, weights=None, use_sh=True, use_cov=False, device="cuda", sh_degree=3):
    """The reference's own CUDA kernels (oracle/_ref) on the same inputs."""
    from oracle.refshim import RefRasterizer
    g = g.to(device)
    cam = cam.to(device)
    bg = bg.to(device)
    S = 0 if g.get_semantics is None else g.get_semantics.shape[1]
    rr = RefRasterizer(S)
    kw = dict(means3D=g.get_xyz.contiguous(), opacities=g.get_opacity.contiguous(), W=cam.image_width,
              H=cam.image_height, viewmatrix=cam.world_view_transform.contiguous(),
              projmatrix=cam.full_proj_transform.contiguous(), campos=cam.camera_center.contiguous(),
              tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), bg=bg,
              semantics=None if S == 0 else g.get_semantics.contiguous(), sh_degree=sh_degree)
    if use_sh:
        kw["shs"] = g.get_features.contiguous()
    else:
        kw["colors_precomp"] = torch.sigmoid(g.get_features[:, 0, :]).contiguous()
    if use_cov:
        kw["cov3D_precomp"] = g.get_covariance(1.0).contiguous()
    else:
        kw["scales"], kw["rotations"] = g.get_scaling.contiguous(), g.get_rotation.contiguous()
    out = rr.forward(**kw)
    out["num_rendered"] = rr.num_rendered
    if weights is not None:
        w = {k: v.to(device).contiguous() for k, v in weights.items()}
        out["grads"] = rr.backward(w["render"], w["semantics"], w["depth"], w["alpha"])
    torch.cuda.synchronize()
    out["_rr"] = rr
    return out


def to_np(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x
    return x.detach().cpu().numpy()


def image_report(a: dict, b: dict, keys=("color", "semantics", "depth", "alpha")):
    """Per-buffer L-inf and the number of pixels above IMG_TOL (threshold-flip candidates)."""
    rep = {}
    for k in keys:
        x, y = to_np(a[k]), to_np(b[k])
        if x.size == 0:
            rep[k] = dict(linf=0.0, n_bad=0, n=0)
            continue
        d = np.abs(x.astype(np.float64) - y.astype(np.float64))
        bad_pix = (d > IMG_TOL).reshape(d.shape[0], -1).any(axis=0)
        rep[k] = dict(linf=float(d.max()), n_bad=int(bad_pix.sum()), n=int(bad_pix.size))
    return rep


def grad_report(a: dict, b: dict, keys=GRAD_KEYS):
    """max|a-b| / max|b| per gradient tensor (the north-star 1e-3 relative criterion)."""
    rep = {}
    for k in keys:
        if k not in a or k not in b or a[k] is None or b[k] is None:
            continue
        x, y = to_np(a[k]).astype(np.float64), to_np(b[k]).astype(np.float64)
        y = y.reshape(x.shape)
        if x.size == 0:
            continue
        scale = max(np.abs(y).max(), 1e-30)
        rep[k] = dict(rel=float(np.abs(x - y).max() / scale), scale=float(scale))
    return rep


def assert_images_close(a, b, max_bad_frac=0.0, what=""):
    rep = image_report(a, b)
    for k, r in rep.items():
        if r["n"] == 0:
            continue
        assert r["n_bad"] <= max_bad_frac * r["n"], f"{what} {k}: L-inf {r['linf']:.3e}, {r['n_bad']}/{r['n']} pixels > {IMG_TOL}"
    return rep


def assert_grads_close(a, b, rtol=GRAD_RTOL, what="", keys=GRAD_KEYS):
    rep = grad_report(a, b, keys)
    assert rep, "no gradients compared"
    for k, r in rep.items():
        assert r["rel"] <= rtol, f"{what} {k}: max|diff|/max|ref| = {r['rel']:.3e} > {rtol}"
    return rep


# ---- mask goldens (tests/golden/mask_*.npz, cut from the reference's own gui/main.py by make_mask_golden.py) ----
def mask_golden_args(z):
    """Keyword arguments for oracle.mask / SemanticHyperplane from one golden file: APE uses the text hyperplane
    + log_scale + threshold, OSH the LinearSVM weight/bias with the reference's fixed 0.5 threshold."""
    osh = int(z["meta"][4]) == 1
    if osh:
        return dict(w=z["svm_weight"].reshape(-1), mode=1, hyperplane_b=float(z["svm_bias"][0]), log_scale=0.0,
                    thresh=0.5)
    return dict(w=z["text"].reshape(-1), mode=0, hyperplane_b=0.0, log_scale=float(z["log_scale"]),
                thresh=float(z["thresh"]))


def assert_mask_matches_golden(got_sim, got_bg, got_idx, z, what, gap_eps=1e-5):
    """Exact codebook row / mask except where the reference's own top-2 logit gap is below gap_eps (arg-max
    near-ties: summation-order effects); sim within the north star's 1e-4.  Returns the exemption count."""
    near = z["top2_gap"] <= gap_eps
    clear = ~near
    assert near.mean() < 1e-2, f"{what}: {near.sum()} near-ties of {near.size}"
    assert np.array_equal(np.asarray(got_idx)[clear], z["idx"][clear]), f"{what}: codebook rows differ"
    d = np.abs(np.asarray(got_sim, np.float64)[clear] - z["sim"].astype(np.float64)[clear])
    # a sim within 1e-6 of the threshold may land on either side (the output is then 0 or ~thresh): exempt + count
    flips = np.asarray(got_bg)[clear] != z["bg_mask"][clear]
    assert flips.sum() <= 2, f"{what}: {flips.sum()} threshold flips"
    assert d[~flips].max() <= IMG_TOL, f"{what}: sim L-inf {d[~flips].max():.3e}"
    return dict(near_ties=int(near.sum()), threshold_flips=int(flips.sum()), linf=float(d[~flips].max()))


def torch_reference_similarity(x, mlp_weight, mlp_bias, lut, w, log_scale=0.0, thresh=0.86, osh_bias=None):
    """The reference's torch op chain restated for the GPU tests (gui/main.py:365-384); the pinned comparison is
    against tests/golden/mask_*.npz, which the reference's own source produced."""
    dec = torch.nn.functional.linear(x, mlp_weight, mlp_bias)
    idx = torch.softmax(dec * 10, dim=-1).argmax(dim=-1)
    f = lut[idx]
    f = f / f.norm(dim=-1, keepdim=True)
    if osh_bias is not None:
        sim = torch.nn.functional.linear(f / 0.3438, w.reshape(1, -1), torch.tensor([osh_bias], device=x.device))
        sim = sim.squeeze().sigmoid()
        thresh = 0.5
    else:
        logit = torch.matmul(f, w.reshape(1, -1).transpose(-1, -2)) / math.exp(log_scale)
        logit = torch.clamp(torch.clamp(logit, max=50000), min=-50000) + 2
        sim = logit.sigmoid().squeeze(-1)
    bg = sim < thresh
    sim = sim.clone()
    sim[bg] = 0
    return sim, bg, idx
