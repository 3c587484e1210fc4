"""CPU tests: the oracle (oracle/goi_oracle.c) against the golden vectors.

* tests/golden/ref_*.npz were produced by the REFERENCE's own CUDA kernels on a B200
  (tests/golden/make_reference_golden.py) -- this is what pins the oracle.
* tests/golden/twin_vectors.npz comes from the reference's Python twins (eval_sh,
  build_covariance_from_scaling_rotation, getProjectionMatrix; tests/golden/make_twin_vectors.py).
"""
import glob
import math
import os

import numpy as np
import pytest
import torch

import common
from goi_b200.scenes import SyntheticCamera, SyntheticGaussians, make_loss_weights, make_scene, projection_matrix
from oracle import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_FILES = sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz")))


def test_golden_files_present():
    assert len(REF_FILES) >= 4 and os.path.exists(os.path.join(GOLDEN, "twin_vectors.npz"))


@pytest.mark.parametrize("path", REF_FILES, ids=[os.path.basename(p) for p in REF_FILES])
def test_oracle_matches_reference_cuda_golden(path):
    z = np.load(path)
    P, W, H, S, seed, use_sh, use_cov = [int(v) for v in z["meta"]]
    g, cam, _ = make_scene(P, W, H, S, seed)
    bg = torch.tensor(z["bg"])
    w = make_loss_weights(S, W, H, seed)
    ora = common.run_oracle(g, cam, bg, w, use_sh=bool(use_sh), use_cov=bool(use_cov), wide=True)
    assert np.array_equal(ora["radii"], z["radii"])
    assert ora["num_rendered"] == int(z["num_rendered"][0])
    ref = {k: z[k] for k in ("color", "semantics", "depth", "alpha")}
    common.assert_images_close(ora, ref, max_bad_frac=5e-4, what="oracle vs reference-CUDA golden")
    keys = ["dL_dmeans3D", "dL_dmeans2D", "dL_dsemantics", "dL_dopacity", "dL_dcolors", "dL_dconic", "dL_ddepths",
            "dL_dcov3D"]
    keys += ["dL_dsh"] if use_sh else []
    keys += [] if use_cov else ["dL_dscales", "dL_drotations"]
    refg = {k: z[k] for k in keys}
    rep = common.assert_grads_close(ora["grads"], refg, rtol=2e-3, what="oracle vs reference-CUDA golden", keys=keys)
    assert len(rep) == len(keys)


def test_float_accumulation_matches_wide_within_tolerance():
    """The reference accumulates per-Gaussian gradients with float atomics in arbitrary order; the oracle's
    float-ordered and double accumulations must agree far inside the 1e-3 gradient tolerance."""
    g, cam, bg = make_scene(800, 96, 64, 10, 5)
    w = make_loss_weights(10, 96, 64, 5)
    a = common.run_oracle(g, cam, bg, w, wide=True)["grads"]
    b = common.run_oracle(g, cam, bg, w, wide=False)["grads"]
    rep = common.grad_report(a, b)
    assert rep and max(r["rel"] for r in rep.values()) < 1e-4


def test_sh_and_cov3d_against_python_twins():
    z = np.load(os.path.join(GOLDEN, "twin_vectors.npz"))
    N = z["pos"].shape[0]
    # a camera at campos looking down +z so that every point at z - campos.z > 0.2 is processed; we only read
    # the per-Gaussian state (rgb, cov3D), which does not depend on the projection
    pos = z["pos"].copy()
    pos[:, 2] = np.abs(pos[:, 2]) + 1.0
    campos = z["campos"]
    w2v = torch.eye(4)
    w2v[:3, 3] = -torch.tensor(campos)
    cam = SyntheticCamera(64, 64, math.radians(90), w2v)
    assert np.allclose(cam.camera_center.numpy(), campos, atol=1e-6)
    d = pos - campos
    dirs = d / np.linalg.norm(d, axis=1, keepdims=True)
    import importlib
    gr = importlib.import_module("gaussian_renderer")
    for deg in range(4):
        res = oracle.forward(means3D=pos, opacities=np.full((N, 1), 0.5, np.float32), shs=z["sh"],
                             scales=np.full((N, 3), 0.05, np.float32), rotations=z["rotations"], sh_degree=deg,
                             **common.cam_arrays(cam, torch.zeros(3)))
        st = res.state()
        vis = res.radii > 0
        assert vis.sum() > N // 4
        # eval_sh twin evaluated at OUR directions (the golden used other positions); check the shared formula
        twin = gr.eval_sh(deg, torch.tensor(z["sh"]).transpose(1, 2), torch.tensor(dirs, dtype=torch.float32))
        twin = torch.clamp_min(twin + 0.5, 0.0).numpy()
        assert np.abs(st["rgb"][vis] - twin[vis]).max() < 2e-6
    # ... and our eval_sh port against the reference's eval_sh outputs on the golden directions
    dg = z["pos"] - campos
    dg = dg / np.linalg.norm(dg, axis=1, keepdims=True)
    for deg in range(4):
        ours = gr.eval_sh(deg, torch.tensor(z["sh"]).transpose(1, 2), torch.tensor(dg, dtype=torch.float32))
        ours = torch.clamp_min(ours + 0.5, 0.0).numpy()
        assert np.abs(ours - z[f"sh2rgb_deg{deg}"]).max() < 2e-6
    # covariance: the oracle's computeCov3D vs build_covariance_from_scaling_rotation (unit quaternions)
    for mod in (1.0, 0.6):
        res = oracle.forward(means3D=pos, opacities=np.full((N, 1), 0.5, np.float32), shs=z["sh"],
                             scales=z["scales"], rotations=z["rotations"], scale_modifier=mod,
                             **common.cam_arrays(cam, torch.zeros(3)))
        st = res.state()
        vis = res.radii > 0
        ref = z[f"cov3D_mod{mod}"]
        assert np.abs(st["cov3D"][vis] - ref[vis]).max() <= 2e-5 * max(1.0, np.abs(ref).max())
        # the duck-typed container's get_covariance is the same twin
        gs = SyntheticGaussians(torch.tensor(pos), None, torch.tensor(z["scales"]), torch.tensor(z["rotations"]), None, None)
        assert np.abs(gs.get_covariance(mod).numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


def test_camera_conventions_against_reference_graphics_utils():
    z = np.load(os.path.join(GOLDEN, "twin_vectors.npz"))
    P = projection_matrix(0.01, 100.0, 1.0471975512, 0.7).numpy()
    assert np.abs(P - z["proj"]).max() < 1e-6
    Rt = np.zeros((4, 4), np.float32)
    Rt[:3, :3] = z["w2v_R"].T
    Rt[:3, 3] = z["w2v_t"]
    Rt[3, 3] = 1
    assert np.abs(Rt - z["w2v"]).max() < 1e-6           # getWorld2View2 with zero translate / unit scale


def test_oracle_backward_finite_differences():
    """Independent sanity check of the restated backward: central differences of the oracle's own forward."""
    P, W, H, S = 6, 32, 32, 3
    g, cam, bg = make_scene(P, W, H, S, 21, px_sigma=4.0)
    g._xyz[:, :2] *= 0.3
    g._opacity[:] = 0.4 + 0.05 * torch.arange(P).float().unsqueeze(1)
    w = {k: v.double().numpy() for k, v in make_loss_weights(S, W, H, 21).items()}
    ca = common.cam_arrays(cam, torch.tensor([0.1, 0.2, 0.3]))

    def loss(arrs):
        r = oracle.forward(**arrs, **ca)
        L = (r.color * w["render"]).sum() + (r.semantics * w["semantics"]).sum() + (r.depth * w["depth"]).sum() \
            + (r.alpha * w["alpha"]).sum()
        return L, r

    base = common.gaussian_arrays(g)
    L0, r0 = loss(base)
    gr = oracle.backward(r0, *[w[k].astype(np.float32) for k in ("render", "semantics", "depth", "alpha")], wide=True)
    checks = [("semantics", "dL_dsemantics", 1e-2), ("opacities", "dL_dopacity", 1e-3), ("means3D", "dL_dmeans3D", 2e-4),
              ("scales", "dL_dscales", 1e-4), ("shs", "dL_dsh", 1e-2)]
    for name, gname, eps in checks:
        arr = base[name]
        flat_idx = [0, arr.size // 2, arr.size - 1]
        for fi in flat_idx:
            ap = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in base.items()}
            am = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in base.items()}
            ap[name].reshape(-1)[fi] += eps
            am[name].reshape(-1)[fi] -= eps
            fd = (loss(ap)[0] - loss(am)[0]) / (2 * eps)
            an = gr[gname].reshape(-1)[fi]
            scale = max(np.abs(gr[gname]).max(), 1e-6)
            assert abs(fd - an) <= 0.05 * scale + 5e-3, (name, fi, fd, an)


def test_oracle_trace_and_mark_visible():
    g, cam, bg = make_scene(300, 48, 32, 4, 9)
    g._xyz[::5, 2] *= -1
    vis = oracle.mark_visible(g.get_xyz.numpy(), cam.world_view_transform.numpy(), cam.full_proj_transform.numpy())
    assert vis.sum() == (g.get_xyz[:, 2] > 0.2).sum().item()
    img = np.random.default_rng(0).random((4, 32, 48)).astype(np.float32)
    ca = common.cam_arrays(cam, bg)
    kw = dict(means3D=g.get_xyz.numpy(), opacities=g.get_opacity.numpy(), shs=g.get_features.numpy(),
              scales=g.get_scaling.numpy(), rotations=g.get_rotation.numpy(), img_sem=img, **ca)
    a = oracle.trace(count_per_channel=True, **kw)
    b = oracle.trace(count_per_channel=False, **kw)
    assert np.array_equal(a["num_gsem"], 4 * b["num_gsem"])      # the reference bumps the counter once per channel
    assert np.allclose(a["gau_sem"], b["gau_sem"])
    fwd = oracle.forward(**common.gaussian_arrays(g), **ca)
    assert np.abs(fwd.color - a["color"]).max() < 1e-6           # same composite for the colour image


# ---------------------------------------------------------------------------------------------
# mask path: the oracle against goldens cut from the reference's own gui/main.py / vision_language_align.py /
# networks.py / semantic_model.py (tests/golden/make_mask_golden.py)
# ---------------------------------------------------------------------------------------------
MASK_FILES = sorted(glob.glob(os.path.join(GOLDEN, "mask_*.npz")))


def test_mask_golden_files_present():
    assert len(MASK_FILES) >= 4


@pytest.mark.parametrize("path", MASK_FILES, ids=[os.path.basename(p) for p in MASK_FILES])
def test_oracle_mask_matches_reference_golden(path):
    z = np.load(path)
    kw = common.mask_golden_args(z)
    o = oracle.mask(z["x"], z["mlp_weight"], z["mlp_bias"], z["lut"], kw.pop("w"), **kw)
    rep = common.assert_mask_matches_golden(o["sim"], o["bg_mask"], o["idx"], z, "oracle mask vs reference golden")
    # the oracle's own near-tie measure must agree with the reference's logits
    assert np.abs(o["top2_gap"] - z["top2_gap"]).max() <= 1e-5
    print(rep)
