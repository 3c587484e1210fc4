"""GPU tests: the CUDA path against the committed reference-CUDA golden vectors, and the fused mask kernel
against the oracle / the reference's torch op chain."""
import glob
import math
import os

import numpy as np
import pytest
import torch

import common
from goi_b200.scenes import make_loss_weights, make_mask_model, make_scene
from common import torch_reference_similarity
from goi_b200.semantic_mask import SemanticHyperplane
from oracle import oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_FILES = sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz")))


@pytest.mark.parametrize("path", REF_FILES, ids=[os.path.basename(p) for p in REF_FILES])
def test_cuda_matches_reference_cuda_golden(path):
    z = np.load(path)
    P, W, H, S, seed, use_sh, use_cov = [int(v) for v in z["meta"]]
    g, cam, _ = make_scene(P, W, H, S, seed)
    bg = torch.tensor(z["bg"])
    w = make_loss_weights(S, W, H, seed)
    cu = common.run_cuda(g, cam, bg, w, use_sh=bool(use_sh), use_cov=bool(use_cov))
    assert np.array_equal(common.to_np(cu["radii"]), z["radii"])
    common.assert_images_close(cu, {k: z[k] for k in ("color", "semantics", "depth", "alpha")}, max_bad_frac=0.0,
                               what="cuda vs reference golden")
    keys = [k for k in cu["grads"] if cu["grads"][k] is not None and k in z.files]
    assert len(keys) >= 5
    common.assert_grads_close(cu["grads"], {k: z[k] for k in keys}, what="cuda vs reference golden", keys=keys)


MASK_FILES = sorted(glob.glob(os.path.join(GOLDEN, "mask_*.npz")))


@pytest.mark.parametrize("channels_first", [False, True])
@pytest.mark.parametrize("path", MASK_FILES, ids=[os.path.basename(p) for p in MASK_FILES])
def test_mask_kernel_matches_reference_golden(path, channels_first):
    """goi_mask (k_mask_table + k_mask_apply) against vectors the reference's own gui/main.py:363-385 +
    vision_language_align.py:109-122 + networks.py LinearSVM + semantic_model.py produced
    (tests/golden/make_mask_golden.py).  [N,S] row-major = the per-Gaussian 3D selection form, channels_first = the
    planar render output."""
    z = np.load(path)
    assert len(MASK_FILES) >= 4
    kw = common.mask_golden_args(z)
    t = lambda a: torch.tensor(np.ascontiguousarray(a)).cuda()
    hp = SemanticHyperplane(t(z["mlp_weight"]), t(z["mlp_bias"]), t(z["lut"]), t(kw["w"]), log_scale=kw["log_scale"],
                            thresh=kw["thresh"])
    if kw["mode"] == 1:
        hp.enable_osh(bias=kw["hyperplane_b"])
    x = t(z["x"])
    N = x.shape[0]
    xin = x.t().contiguous() if channels_first else x
    bg = torch.zeros(N, dtype=torch.bool, device="cuda")
    sim, idx = hp.compute_similarity(xin, out_bg_mask=bg, channels_first=channels_first, want_idx=True)
    rep = common.assert_mask_matches_golden(sim.cpu().numpy(), bg.cpu().numpy(), idx.cpu().numpy(), z,
                                            "goi_mask vs reference golden")
    print("\n", os.path.basename(path), rep)


@pytest.mark.parametrize("S,N,mode,channels_first", [(16, 200_000, "ape", False), (10, 50_000, "ape", True),
                                                     (32, 120_000, "osh", True), (4, 30_001, "ape", False),
                                                     (64, 20_000, "osh", False)])
def test_mask_kernel_matches_oracle_and_torch_chain(S, N, mode, channels_first):
    gen = torch.Generator().manual_seed(S + N)
    x = torch.randn(N, S, generator=gen)
    mlp_w, mlp_b, lut, w = make_mask_model(S, seed=S)
    hp = SemanticHyperplane(mlp_w.cuda(), mlp_b.cuda(), lut.cuda(), w.cuda(), log_scale=0.2, thresh=0.86)
    kw = dict(mode=0, log_scale=0.2, thresh=0.86)
    if mode == "osh":
        hp.enable_osh()
        kw = dict(mode=1, hyperplane_b=hp.svm_bias, thresh=0.5)
    xin = x.t().contiguous().cuda() if channels_first else x.cuda()
    bg_out = torch.zeros(N, dtype=torch.bool, device="cuda")
    sim, idx = hp.compute_similarity(xin, out_bg_mask=bg_out, channels_first=channels_first, want_idx=True)
    o = oracle.mask(x.numpy(), mlp_w.numpy(), mlp_b.numpy(), lut.numpy(), w.numpy(), **kw)
    clear = o["top2_gap"] > 1e-4
    assert clear.mean() > 0.98
    assert np.array_equal(idx.cpu().numpy()[clear], o["idx"][clear]), "codebook rows differ beyond near-ties"
    assert np.abs(sim.cpu().numpy()[clear] - o["sim"][clear]).max() <= 1e-4
    assert np.array_equal(bg_out.cpu().numpy()[clear], o["bg_mask"][clear])
    # and against the reference's own torch expression chain evaluated on the GPU
    if mode == "ape":
        tsim, tbg, tidx = torch_reference_similarity(x.cuda(), mlp_w.cuda(), mlp_b.cuda(), lut.cuda(), w.cuda(),
                                                     log_scale=0.2, thresh=0.86)
    else:
        tsim, tbg, tidx = torch_reference_similarity(x.cuda(), mlp_w.cuda(), mlp_b.cuda(), lut.cuda(), w.cuda(),
                                                     osh_bias=hp.svm_bias)
    same = (tidx.int() == idx)
    assert float(same.float().mean()) > 0.999            # TF32-free fp32 GEMM vs fused dot: near-tie flips only
    assert float((tsim - sim)[same].abs().max()) <= 1e-4


def test_mask_from_render_and_gaussian_selection():
    P, W, H, S = 20_000, 320, 200, 16
    g, cam, bg = make_scene(P, W, H, S, 41)
    cu = common.run_cuda(g, cam, bg)
    mlp_w, mlp_b, lut, w = make_mask_model(S, seed=41)
    hp = SemanticHyperplane(mlp_w.cuda(), mlp_b.cuda(), lut.cuda(), w.cuda(), thresh=0.86)
    mask = hp.mask_from_render(cu["semantics"])
    assert mask.shape == (H, W) and mask.dtype == torch.bool
    # reference route: permute to [HW,S] then the torch chain (gui/main.py:588, 363-385)
    sem = cu["semantics"].permute(1, 2, 0).reshape(-1, S)
    tsim, _, _ = torch_reference_similarity(sem, mlp_w.cuda(), mlp_b.cuda(), lut.cuda(), w.cuda(), thresh=0.86)
    agree = ((tsim > 0).view(H, W) == mask).float().mean()
    assert float(agree) > 0.999
    sel = hp.select_gaussians(g.get_semantics.cuda())
    assert sel.shape == (P,) and 0 < int(sel.sum()) < P


@pytest.mark.parametrize("S,mode,want_sem", [(16, "ape", True), (16, "ape", False), (10, "osh", True),
                                             (32, "ape", False), (4, "osh", False)])
def test_fused_mask_epilogue_equals_render_then_mask(S, mode, want_sem):
    """goi_forward_mask == goi_forward followed by goi_mask, bit for bit, and against the
    oracle's mask of the oracle-independent rendered features."""
    from gaussian_renderer import render, render_mask
    from goi_b200.scenes import PipeFlags
    P, W, H = 30_000, 333, 207                      # ragged: W, H not multiples of the 16-pixel tile
    g, cam, bg = make_scene(P, W, H, S, 50 + S)
    g, cam, bg = g.to("cuda"), cam.to("cuda"), bg.to("cuda")
    mlp_w, mlp_b, lut, w = make_mask_model(S, seed=S)
    hp = SemanticHyperplane(mlp_w.cuda(), mlp_b.cuda(), lut.cuda(), w.cuda(), log_scale=0.1, thresh=0.86)
    if mode == "osh":
        hp.enable_osh()
    with torch.no_grad():
        base = render(cam, g, PipeFlags(), bg)
        bg_ref = torch.zeros(H * W, dtype=torch.bool, device="cuda")
        sim_ref, idx_ref = hp.compute_similarity(base["semantics"], out_bg_mask=bg_ref, channels_first=True,
                                                 want_idx=True)
    out = render_mask(cam, g, PipeFlags(), bg, hp, want_semantics=want_sem, want_idx=True)
    # The plain render accumulates the payload on the tensor cores (3xTF32), the fused-mask variant with FP32 FMAs:
    # identical blending decisions (alpha, radii bit-equal), payload sums equal to rounding.
    assert torch.equal(out["alpha"], base["alpha"]) and torch.equal(out["radii"], base["radii"])
    for k in ("render", "depth") + (("semantics",) if want_sem else ()):
        assert float((out[k] - base[k]).abs().max()) <= 2e-5 * max(1.0, float(base[k].abs().max())), k
    if not want_sem:
        assert out["semantics"] is None
    # the mask of the fused epilogue (warp-level MMA) = goi_mask (tcgen05) applied to ITS OWN semantic image: both
    # evaluate the 3xTF32 projection, in different summation orders -> equal except at arg-max near-ties
    if want_sem:
        bg_own = torch.zeros(H * W, dtype=torch.bool, device="cuda")
        sim_own, idx_own = hp.compute_similarity(out["semantics"], out_bg_mask=bg_own, channels_first=True, want_idx=True)
        same_own = out["idx"].view(-1) == idx_own
        assert float(same_own.float().mean()) > 0.9995
        assert torch.equal(out["sim"].view(-1)[same_own], sim_own[same_own])
        assert torch.equal(out["bg_mask"].view(-1)[same_own], bg_own[same_own])
    # and equal to the mask of the plain render except at arg-max near-ties (the two semantic images differ by ~1e-6)
    same = out["idx"].view(-1) == idx_ref
    assert float(same.float().mean()) > 0.999
    assert torch.equal(out["sim"].view(-1)[same], sim_ref[same]) and torch.equal(out["bg_mask"].view(-1)[same], bg_ref[same])
    assert out["mask"].shape == (H, W) and 0 < int(out["mask"].sum()) < H * W
    # and against the CPU oracle's mask on the same rendered features
    kw = dict(mode=0, log_scale=0.1, thresh=0.86) if mode == "ape" else dict(mode=1, hyperplane_b=hp.svm_bias, thresh=0.5)
    o = oracle.mask(base["semantics"].permute(1, 2, 0).reshape(-1, S).cpu().numpy(), mlp_w.numpy(), mlp_b.numpy(),
                    lut.numpy(), w.numpy(), **kw)
    clear = o["top2_gap"] > 1e-4
    assert np.array_equal(out["idx"].view(-1).cpu().numpy()[clear], o["idx"][clear])
    assert np.abs(out["sim"].view(-1).cpu().numpy()[clear] - o["sim"][clear]).max() <= 1e-4


def test_fused_mask_on_an_empty_scene():
    """P == 0: nothing is rendered; every pixel's feature vector is zero, so the logits are the biases."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    S, W, H = 8, 64, 48
    _, cam, bg = make_scene(10, W, H, S, 1)
    cam, bg = cam.to("cuda"), bg.to("cuda")
    mlp_w, mlp_b, lut, w = make_mask_model(S, seed=3)
    hp = SemanticHyperplane(mlp_w.cuda(), mlp_b.cuda(), lut.cuda(), w.cuda(), thresh=0.5)
    rs = GaussianRasterizationSettings(H, W, math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2), bg, 1.0,
                                       cam.world_view_transform, cam.full_proj_transform, 3, cam.camera_center,
                                       False, False)
    z = lambda *s: torch.zeros(*s, device="cuda")
    out = GaussianRasterizer(rs).forward_mask(z(0, 3), z(0, 1), hp, shs=z(0, 16, 3), semantics=z(0, S),
                                              scales=z(0, 3), rotations=z(0, 4), want_idx=True)
    sim_ref, idx_ref = hp.compute_similarity(z(H * W, S), want_idx=True)
    assert torch.equal(out["idx"].view(-1), idx_ref) and torch.equal(out["sim"].view(-1), sim_ref)
    assert float(out["semantics"].abs().max()) == 0.0 and float(out["render"].abs().max()) == 0.0
