"""GPU parity tests (run on the B200 box: pytest -m gpu).

Every test goes through the reference-shaped Python surface -> C ABI -> CUDA kernels and compares with
(a) the CPU oracle on the same seeded inputs, (b) the reference's own CUDA core (oracle/_ref) when the
prebuilt library travelled with the snapshot, (c) size-independent properties at BASELINE.json's full
config-2 size.  Tolerances are the north star's: 1e-4 L-inf on images, 1e-3 of max on gradients.
"""
import math

import numpy as np
import pytest
import torch

import common
from common import (assert_grads_close, assert_images_close, grad_report, image_report, run_cuda, run_oracle,
                    run_reference_cuda, to_np)
from goi_b200.scenes import SyntheticCamera, SyntheticGaussians, make_loss_weights, make_scene

pytestmark = pytest.mark.gpu
GRAD_RTOL_CHAIN = common.GRAD_RTOL


def _yaw(deg):
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return torch.tensor([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1.0]])


def _ref_ok(S):
    from oracle import refshim
    return refshim.available(S)


# ---------------------------------------------------------------------------------------------
# (a) CUDA vs CPU oracle, config-1-sized cases (10k Gaussians, 256x256) + ragged / option variants
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("P,W,H,S,use_sh,use_cov,seed", [
    (10_000, 256, 256, 0, True, False, 0),      # BASELINE config 1: RGB only
    (10_000, 256, 256, 10, True, False, 0),     # the reference's shipped channel count
    (10_000, 256, 256, 16, True, False, 1),
    (6_000, 250, 197, 16, False, False, 2),     # ragged image, precomputed colours
    (6_000, 200, 120, 4, True, True, 3),        # precomputed 3D covariance
    (4_000, 160, 96, 32, True, False, 4),
    (3_000, 128, 80, 64, True, False, 5),       # widest supported vector (reference cannot build S=64)
    (3_000, 128, 80, 7, True, False, 6),        # S not a multiple of 4: scalar staging path
    (3_000, 128, 80, 20, True, False, 7),       # S between kernel instantiations (padded lanes)
])
def test_cuda_matches_oracle(P, W, H, S, use_sh, use_cov, seed):
    g, cam, bg = make_scene(P, W, H, S, seed)
    bg = torch.tensor([0.3, 0.5, 0.7])
    w = make_loss_weights(S, W, H, seed)
    ora = run_oracle(g, cam, bg, w, use_sh, use_cov)
    cu = run_cuda(g, cam, bg, w, use_sh, use_cov)
    assert np.array_equal(to_np(cu["radii"]), ora["radii"]), "radii differ"
    # threshold flips (1-ulp expf / FMA differences between CPU and GPU) may touch isolated pixels
    rep = assert_images_close(cu, ora, max_bad_frac=2e-4, what="cuda vs oracle")
    keys = [k for k in common.GRAD_KEYS if k in cu["grads"]] + (["dL_dcolors"] if not use_sh else []) \
        + (["dL_dcov3D"] if use_cov else [])
    grep = assert_grads_close(cu["grads"], ora["grads"], rtol=2e-3, what="cuda vs oracle", keys=keys)
    print("\n", rep, "\n", grep)


# ---------------------------------------------------------------------------------------------
# (b) CUDA vs the reference's own CUDA kernels (same compiler, same expression order)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("P,W,H,S,seed", [
    (10_000, 256, 256, 0, 0),
    (10_000, 256, 256, 10, 0),
    (200_000, 800, 600, 16, 1),
    (100_000, 640, 360, 32, 2),
])
def test_cuda_matches_reference_cuda(P, W, H, S, seed):
    if not _ref_ok(S):
        pytest.skip("oracle/_ref not built")
    g, cam, bg = make_scene(P, W, H, S, seed)
    bg = torch.tensor([0.1, 0.2, 0.3])
    w = make_loss_weights(S, W, H, seed)
    ref = run_reference_cuda(g, cam, bg, w)
    cu = run_cuda(g, cam, bg, w)
    assert torch.equal(cu["radii"].cpu(), ref["radii"].cpu()), "radii differ from the reference kernels"
    rep = assert_images_close(cu, ref, max_bad_frac=0.0, what="cuda vs reference cuda")
    grep = assert_grads_close(cu["grads"], ref["grads"], what="cuda vs reference cuda")
    print("\n", rep, "\n", grep)


@pytest.mark.parametrize("n_big,factor", [(64, 30.0), (500, 8.0), (5, 400.0)])
def test_large_splats_match_reference_cuda(n_big, factor):
    """Splats whose tile rectangle exceeds 32 tiles take the warp-cooperative paths (k_count_big_rects,
    cooperative emission): images, gradients and radii must still equal the reference kernels', and two
    runs must agree on the instance count (count and emission use the same test)."""
    S, P, W, H, seed = 16, 30_000, 500, 333, 7
    if not _ref_ok(S):
        pytest.skip("oracle/_ref not built")
    g, cam, bg = make_scene(P, W, H, S, seed)
    with torch.no_grad():
        g._scaling[:n_big] *= factor                 # consecutive indices: the worst case for a per-thread loop
        g._opacity[:n_big] = g._opacity[:n_big] * 0.1 + 0.01
    bg = torch.tensor([0.2, 0.1, 0.4])
    w = make_loss_weights(S, W, H, seed)
    ref = run_reference_cuda(g, cam, bg, w)
    cu = run_cuda(g, cam, bg, w)
    assert torch.equal(cu["radii"].cpu(), ref["radii"].cpu())
    assert int((cu["radii"] > 64).sum()) >= min(n_big, 5) // 2      # the large ones are really on screen
    assert_images_close(cu, ref, max_bad_frac=0.0, what="large splats vs reference cuda")
    assert_grads_close(cu["grads"], ref["grads"], what="large splats vs reference cuda")
    from diff_gaussian_rasterization import _C
    r1 = _C.num_rendered()
    run_cuda(g, cam, bg)
    assert _C.num_rendered() == r1


@pytest.mark.parametrize("aspect,angle_deg", [(200.0, 45.0), (1000.0, 45.0), (300.0, 30.0), (500.0, 135.0)])
def test_needle_splats_match_reference_cuda(aspect, angle_deg):
    """Very elongated diagonal splats: the three products of `power` reach 1e4..1e6 and cancel, so float rounding of
    the bound in goi_cull.cuh is proportional to their magnitude (its slack scales with it).  A pair the reference
    blends at alpha ~ 1/255 must not be culled at tile or warp level: the images and radii must equal the reference
    kernels' with zero bad pixels, and so must every gradient the composite produces (mean2D, opacity, semantics, SH).
    The chain from dL/dconic through the 2D covariance to means3D / scales / rotations is ill-conditioned for needles
    (det^2 division, backward.cu:186-204: summation-order noise of 1e-7 in dL/dconic is amplified ~1e5), so there the
    reference's own float-atomic result is itself noise at the 1e-2 level; those three are held to
    `not worse than the reference` against the CPU oracle's double-precision accumulation."""
    S, P, W, H, seed = 16, 6_000, 480, 300, 13
    if not _ref_ok(S):
        pytest.skip("oracle/_ref not built")
    g, cam, bg = make_scene(P, W, H, S, seed)
    n = 1500
    with torch.no_grad():
        # needles in the image plane: long axis x, rotated about the view axis z by angle_deg
        a = math.radians(angle_deg) * (1.0 + 0.02 * torch.randn(n, generator=torch.Generator().manual_seed(1)))
        g._rotation[:n] = torch.stack([torch.cos(a / 2), torch.zeros(n), torch.zeros(n), torch.sin(a / 2)], dim=1)
        s = g._scaling[:n, 1:2] * 0.3
        g._scaling[:n] = torch.cat([s * aspect, s, s], dim=1)
        g._opacity[:n] = 0.02 + 0.1 * g._opacity[:n]              # faint: the 1/255 cut runs through the splat
    bg = torch.tensor([0.2, 0.1, 0.4])
    w = make_loss_weights(S, W, H, seed)
    ref = run_reference_cuda(g, cam, bg, w)
    cu = run_cuda(g, cam, bg, w)
    assert torch.equal(cu["radii"].cpu(), ref["radii"].cpu())
    assert int((cu["radii"][:n] > 200).sum()) > n // 10           # the needles really are hundreds of pixels long
    assert_images_close(cu, ref, max_bad_frac=0.0, what="needle splats vs reference cuda")
    strict = ("dL_dmeans2D", "dL_dopacity", "dL_dsemantics", "dL_dsh")
    assert_grads_close(cu["grads"], ref["grads"], what="needle splats vs reference cuda", keys=strict)
    ora = run_oracle(g, cam, bg, w, wide=True)["grads"]
    chain = ("dL_dmeans3D", "dL_dscales", "dL_drotations")
    ours, theirs = grad_report(cu["grads"], ora, chain), grad_report(ref["grads"], ora, chain)
    for k in chain:
        assert ours[k]["rel"] <= max(GRAD_RTOL_CHAIN, 3.0 * theirs[k]["rel"]), \
            f"{k}: ours {ours[k]['rel']:.2e} vs the reference's own {theirs[k]['rel']:.2e} (both against the wide oracle)"
    print("\nneedles", aspect, angle_deg, {k: (f"{ours[k]['rel']:.1e}", f"{theirs[k]['rel']:.1e}") for k in chain})


def test_dense_skewed_scene_matches_reference_cuda():
    """Most splats squeezed into the bottom rows and made opaque: very uneven tile lists (exercises the
    longest-list-first block order), pixels that saturate early (forward termination, the backward's per-warp
    bound on the walk), against the reference kernels."""
    S, P, W, H, seed = 16, 120_000, 480, 320, 9
    if not _ref_ok(S):
        pytest.skip("oracle/_ref not built")
    g, cam, bg = make_scene(P, W, H, S, seed)
    with torch.no_grad():
        n = int(0.8 * P)
        z = g._xyz[:n, 2]
        ymax = z * (H / W) * math.tan(math.radians(30.0))
        g._xyz[:n, 1] = ymax * (1.0 - 0.1 * torch.rand(n, generator=torch.Generator().manual_seed(2)))
        g._opacity[:n] = 0.6 + 0.39 * g._opacity[:n]
    bg = torch.tensor([0.0, 0.3, 0.1])
    w = make_loss_weights(S, W, H, seed)
    ref = run_reference_cuda(g, cam, bg, w)
    cu = run_cuda(g, cam, bg, w)
    assert torch.equal(cu["radii"].cpu(), ref["radii"].cpu())
    assert float((to_np(ref["alpha"]) > 0.999).mean()) > 0.05          # a good part of the image really saturates
    assert_images_close(cu, ref, max_bad_frac=0.0, what="dense skewed scene vs reference cuda")
    assert_grads_close(cu["grads"], ref["grads"], what="dense skewed scene vs reference cuda")


@pytest.mark.parametrize("P,W,H,S,seed", [(10_000, 256, 256, 10, 0), (8_000, 250, 197, 16, 3)])
def test_oracle_pinned_by_reference_cuda(P, W, H, S, seed):
    """The CPU oracle itself is checked against the real reference kernels (parity pin)."""
    if not _ref_ok(S):
        pytest.skip("oracle/_ref not built")
    g, cam, bg = make_scene(P, W, H, S, seed)
    bg = torch.tensor([0.3, 0.5, 0.7])
    w = make_loss_weights(S, W, H, seed)
    ref = run_reference_cuda(g, cam, bg, w)
    ora = run_oracle(g, cam, bg, w, wide=True)
    assert np.array_equal(to_np(ref["radii"]), ora["radii"])
    assert ref["num_rendered"] == ora["num_rendered"]
    assert_images_close(ora, ref, max_bad_frac=2e-4, what="oracle vs reference cuda")
    keys = list(common.GRAD_KEYS) + ["dL_dcolors", "dL_dconic", "dL_ddepths", "dL_dcov3D"]
    assert_grads_close(ora["grads"], ref["grads"], rtol=2e-3, what="oracle vs reference cuda", keys=keys)


# ---------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------
def _settings(cam, bg, device="cuda"):
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    cam = cam.to(device)
    return GaussianRasterizationSettings(cam.image_height, cam.image_width, math.tan(cam.FoVx / 2),
                                         math.tan(cam.FoVy / 2), bg.to(device), 1.0, cam.world_view_transform,
                                         cam.full_proj_transform, 3, cam.camera_center, False, False)


def test_empty_scene():
    from diff_gaussian_rasterization import GaussianRasterizer
    cam = SyntheticCamera(64, 48, math.radians(60))
    rast = GaussianRasterizer(_settings(cam, torch.tensor([0.2, 0.4, 0.6])))
    z = lambda *s: torch.zeros(*s, device="cuda")
    color, sem, radii, depth, alpha = rast(means3D=z(0, 3), means2D=z(0, 3), opacities=z(0, 1), shs=z(0, 16, 3),
                                           semantics=z(0, 16), scales=z(0, 3), rotations=z(0, 4))
    assert color.shape == (3, 48, 64) and radii.shape == (0,)
    assert float(color.abs().max()) == 0.0 and float(alpha.abs().max()) == 0.0   # reference: zero fill, no bg


def test_all_culled_renders_background():
    from diff_gaussian_rasterization import GaussianRasterizer
    g, cam, _ = make_scene(500, 64, 48, 16, 0)
    g._xyz[:, 2] = -g._xyz[:, 2]                  # everything behind the camera
    bg = torch.tensor([0.2, 0.4, 0.6])
    cu = run_cuda(g, cam, bg, make_loss_weights(16, 64, 48, 0))
    assert int((cu["radii"] > 0).sum()) == 0
    assert torch.allclose(cu["color"], bg.cuda().view(3, 1, 1).expand(3, 48, 64))
    assert float(cu["alpha"].detach().abs().max()) == 0.0 and float(cu["semantics"].detach().abs().max()) == 0.0
    for k, v in cu["grads"].items():
        assert v is None or float(v.abs().max()) == 0.0, k


def test_huge_and_tiny_gaussians():
    """One Gaussian covering every tile + sub-pixel Gaussians + zero-opacity ones."""
    g, cam, bg = make_scene(300, 96, 64, 16, 9)
    g._scaling[0] = torch.tensor([30.0, 30.0, 30.0]); g._xyz[0] = torch.tensor([0.0, 0.0, 5.0])
    g._scaling[1:50] *= 1e-3
    g._opacity[50:80] = 0.0
    g._opacity[80:90] = 1.0
    w = make_loss_weights(16, 96, 64, 9)
    ora = run_oracle(g, cam, bg, w)
    cu = run_cuda(g, cam, bg, w)
    assert np.array_equal(to_np(cu["radii"]), ora["radii"])
    assert_images_close(cu, ora, max_bad_frac=5e-4, what="huge/tiny")
    assert_grads_close(cu["grads"], ora["grads"], rtol=2e-3, what="huge/tiny")


def test_equal_depth_ties_resolve_by_index():
    """Stable sort contract: identical depths composite in ascending Gaussian index."""
    P, W, H, S = 64, 48, 48, 4
    g, cam, bg = make_scene(P, W, H, S, 11)
    g._xyz[:, 2] = 4.0                              # all the same view-space depth
    g._xyz[:, :2] *= 0.2
    g._opacity[:] = 0.9
    ora = run_oracle(g, cam, bg)
    cu = run_cuda(g, cam, bg)
    assert_images_close(cu, ora, max_bad_frac=0.0, what="depth ties")


def test_mark_visible_and_trace():
    from diff_gaussian_rasterization import GaussianRasterizer
    from oracle import oracle
    g, cam, bg = make_scene(2000, 96, 64, 10, 13)
    g._xyz[::7, 2] *= -1
    rast = GaussianRasterizer(_settings(cam, bg))
    vis = rast.markVisible(g.get_xyz.cuda())
    exp = oracle.mark_visible(g.get_xyz.numpy(), cam.world_view_transform.numpy(), cam.full_proj_transform.numpy())
    assert np.array_equal(vis.cpu().numpy(), exp)

    gen = torch.Generator().manual_seed(5)
    img_sem = torch.rand(10, 64, 96, generator=gen)
    gd = g.to("cuda")
    color, gau_sem, num_gsem = rast.trace(means3D=gd.get_xyz, means2D=torch.zeros_like(gd.get_xyz),
                                          opacities=gd.get_opacity, shs=gd.get_features, img_sem=img_sem.cuda(),
                                          scales=gd.get_scaling, rotations=gd.get_rotation)
    ca = common.cam_arrays(cam, bg)
    exp = oracle.trace(means3D=g.get_xyz.numpy(), opacities=g.get_opacity.numpy(), shs=g.get_features.numpy(),
                       scales=g.get_scaling.numpy(), rotations=g.get_rotation.numpy(), img_sem=img_sem.numpy(), **ca)
    assert np.abs(color.cpu().numpy() - exp["color"]).max() <= 1e-4
    # counts are integers: exact except for pairs whose alpha sits on the 0.005 / (1/255) thresholds
    cnt, ecnt = num_gsem.cpu().numpy(), exp["num_gsem"]
    flipped = cnt != ecnt
    assert flipped.mean() < 2e-3
    # rows without a threshold flip: a sum of <= a few hundred float atomics against the oracle's sequential sum
    got, want = gau_sem.cpu().numpy()[~flipped], exp["gau_sem"][~flipped]
    err = np.abs(got - want).max(axis=1)
    tol = 1e-5 * (np.abs(want).max(axis=1) + 1.0)
    assert (err > tol).mean() < 1e-3, f"{(err > tol).sum()} of {err.size} rows off by more than 1e-5 relative"
    # a flipped pair moves a row by one img_sem value (<= 1) per channel
    assert np.abs(gau_sem.cpu().numpy()[flipped] - exp["gau_sem"][flipped]).max(initial=0.0) <= 3.0


# ---------------------------------------------------------------------------------------------
# (b') CUDA vs the reference's own CUDA kernels at the FULL sizes BASELINE.json quotes (configs 2-5)
# ---------------------------------------------------------------------------------------------
FULL_SIZE = [
    ("c2", 1_000_000, 1600, 1000, 16, 1),          # the headline metric's config
    ("c3", 1_000_000, 800, 600, 32, 2),
    ("c5_4", 1_000_000, 1280, 720, 4, 4),
    ("c5_8", 1_000_000, 1280, 720, 8, 4),
    ("c5_16", 1_000_000, 1280, 720, 16, 4),
    ("c5_32", 1_000_000, 1280, 720, 32, 4),
]


@pytest.mark.parametrize("name,P,W,H,S,seed", FULL_SIZE, ids=[c[0] for c in FULL_SIZE])
def test_full_size_cuda_matches_reference_cuda(name, P, W, H, S, seed):
    """Same bounds as test_cuda_matches_reference_cuda -- radii equal, images <= 1e-4 with ZERO bad pixels, gradients
    <= 1e-3 of the tensor's max -- on the scenes bench.py times (same generator, same seed, same first camera)."""
    if not _ref_ok(S):
        pytest.skip("oracle/_ref not built")
    g, cam, bg = make_scene(P, W, H, S, seed)
    bg = torch.tensor([0.1, 0.2, 0.3])
    w = make_loss_weights(S, W, H, seed)
    ref = run_reference_cuda(g, cam, bg, w)
    cu = run_cuda(g, cam, bg, w)
    assert torch.equal(cu["radii"].cpu(), ref["radii"].cpu()), "radii differ from the reference kernels"
    rep = assert_images_close(cu, ref, max_bad_frac=0.0, what=f"{name}: cuda vs reference cuda")
    grep = assert_grads_close(cu["grads"], ref["grads"], what=f"{name}: cuda vs reference cuda")
    from diff_gaussian_rasterization import _C
    print("\n", name, "instances: ours", _C.num_rendered(), "reference", ref["num_rendered"], "\n",
          {k: f"{v['linf']:.2e}" for k, v in rep.items()}, "\n", {k: f"{v['rel']:.2e}" for k, v in grep.items()})


def test_full_size_c4_view_matches_reference_cuda():
    """BASELINE config 4: 5M Gaussians, 1920x1080, S=16, orbit cameras -- one of the 64 views against the reference."""
    from goi_b200.scenes import make_orbit_scene
    P, W, H, S = 5_000_000, 1920, 1080, 16
    if not _ref_ok(S):
        pytest.skip("oracle/_ref not built")
    g, cams, bg = make_orbit_scene(P, W, H, S, 64, 3)
    w = make_loss_weights(S, W, H, 3)
    cam = cams[5]
    ref = run_reference_cuda(g, cam, bg, w)
    cu = run_cuda(g, cam, bg, w)
    assert torch.equal(cu["radii"].cpu(), ref["radii"].cpu())
    assert_images_close(cu, ref, max_bad_frac=0.0, what="c4 view vs reference cuda")
    assert_grads_close(cu["grads"], ref["grads"], what="c4 view vs reference cuda")


def test_full_size_c5_64_matches_oracle():
    """S=64 at the config-5 size: the reference does not build at this width (its backward needs 76.8 KB of static
    shared memory), so the check is against the CPU oracle: threshold flips counted."""
    P, W, H, S, seed = 1_000_000, 1280, 720, 64, 4
    g, cam, bg = make_scene(P, W, H, S, seed)
    w = make_loss_weights(S, W, H, seed)
    ora = run_oracle(g, cam, bg, w)
    cu = run_cuda(g, cam, bg, w)
    assert np.array_equal(to_np(cu["radii"]), ora["radii"])
    assert_images_close(cu, ora, max_bad_frac=2e-4, what="c5_64 vs oracle")
    assert_grads_close(cu["grads"], ora["grads"], rtol=2e-3, what="c5_64 vs oracle")


# ---------------------------------------------------------------------------------------------
# the C ABI's other forward entry points: goi_forward (allocator callbacks -- the form INTEGRATION.md section B
# binds) and the two-phase goi_forward_prepare / goi_forward_render
# ---------------------------------------------------------------------------------------------
def test_callback_and_two_phase_forward_equal_forward_auto():
    import ctypes as C
    from diff_gaussian_rasterization import _C
    P, W, H, S = 20_000, 333, 207, 16
    g, cam, bg = make_scene(P, W, H, S, 61)
    g, cam, bg = g.to("cuda"), cam.to("cuda"), torch.tensor([0.3, 0.1, 0.2], device="cuda")
    base = run_cuda(g, cam, bg)
    L = _C.lib()
    keep = []
    gs, _ = _C._make_gaussians(g.get_xyz, g.get_features, None, g.get_semantics, g.get_opacity, g.get_scaling,
                               g.get_rotation, None, keep)
    view = _C._make_view(bg, cam.world_view_transform, cam.full_proj_transform, cam.camera_center, W, H,
                         math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2), 1.0, 3, False, False, keep)
    stream = torch.cuda.current_stream().cuda_stream

    def outputs():
        f = dict(dtype=torch.float32, device="cuda")
        o = dict(color=torch.full((3, H, W), float("nan"), **f), semantics=torch.full((S, H, W), float("nan"), **f),
                 depth=torch.full((1, H, W), float("nan"), **f), alpha=torch.full((1, H, W), float("nan"), **f),
                 radii=torch.full((P,), -7, dtype=torch.int32, device="cuda"))
        return o, _C.goi_fwd_out(o["color"].data_ptr(), o["semantics"].data_ptr(), o["depth"].data_ptr(),
                                 o["alpha"].data_ptr(), o["radii"].data_ptr())

    def same(o):
        for k in ("color", "semantics", "depth", "alpha", "radii"):
            assert torch.equal(o[k], base[k]), k

    # (1) goi_forward: the library asks the caller for its three scratch blobs through callbacks
    sizes = {"g": [], "b": [], "i": []}
    stores = {"g": [], "b": [], "i": []}

    def mk(tag):
        def alloc(_user, nbytes):
            sizes[tag].append(int(nbytes))
            t = torch.empty((max(int(nbytes), 1),), dtype=torch.uint8, device="cuda")
            stores[tag].append(t)
            return t.data_ptr()
        return _C.ALLOC_FN(alloc)
    cbs = {t: mk(t) for t in "gbi"}
    o, fo = outputs()
    R = C.c_int64(0)
    _C._check(L.goi_forward(C.byref(view), C.byref(gs), C.byref(fo), cbs["g"], None, cbs["b"], None, cbs["i"], None,
                            stream, C.byref(R)), "goi_forward")
    torch.cuda.synchronize()
    same(o)
    assert R.value == _C.num_rendered() and R.value > 0
    assert sizes["g"] == [L.goi_geom_bytes(P, S)] and sizes["i"] == [L.goi_image_bytes(W, H)]
    assert sizes["b"] == [L.goi_binning_bytes(R.value)]          # sized from the exact instance count

    # (2) prepare -> (caller sizes the binning blob) -> render
    o, fo = outputs()
    geom = torch.empty((L.goi_geom_bytes(P, S),), dtype=torch.uint8, device="cuda")
    img = torch.empty((L.goi_image_bytes(W, H),), dtype=torch.uint8, device="cuda")
    R2 = C.c_int64(0)
    _C._check(L.goi_forward_prepare(C.byref(view), C.byref(gs), o["radii"].data_ptr(), geom.data_ptr(), geom.numel(),
                                    stream, C.byref(R2)), "goi_forward_prepare")
    assert R2.value == R.value
    binning = torch.empty((L.goi_binning_bytes(R2.value),), dtype=torch.uint8, device="cuda")
    _C._check(L.goi_forward_render(C.byref(view), C.byref(gs), C.byref(fo), geom.data_ptr(), geom.numel(),
                                   binning.data_ptr(), binning.numel(), img.data_ptr(), img.numel(), R2.value, stream),
              "goi_forward_render")
    torch.cuda.synchronize()
    same(o)
    # an undersized binning blob is reported, not overrun
    rc = L.goi_forward_render(C.byref(view), C.byref(gs), C.byref(fo), geom.data_ptr(), geom.numel(),
                              binning.data_ptr(), 1024, img.data_ptr(), img.numel(), R2.value, stream)
    assert rc == -3 and b"binning" in L.goi_last_error()


def test_async_forward_equals_synchronous_and_flags_overflow():
    """goi_forward_async (no num_rendered read-back: capacity-sized binning, padded key tail) must give bit-identical
    images and gradients to the synchronous path, and a view that does not fit its capacity must be flagged."""
    from diff_gaussian_rasterization import _C
    P, W, H, S = 40_000, 400, 267, 16
    g, cam, bg = make_scene(P, W, H, S, 71)
    w = make_loss_weights(S, W, H, 71)
    dev = torch.device("cuda", torch.cuda.current_device())
    _C.device_state(dev).r_guess = 0                      # (the estimate decays slowly: forget earlier tests' scenes)
    base = run_cuda(g, cam, bg, w)                        # synchronous; seeds the instance-count estimate
    R = _C.num_rendered()
    try:
        _C.set_async_binning(True, dev, headroom=1.07)
        a = run_cuda(g, cam, bg, w)
        assert _C.check_async(dev, wait=True) == 1 and _C.num_rendered() == R
        for k in ("color", "semantics", "depth", "alpha", "radii"):
            assert torch.equal(a[k], base[k]), k
        for k, v in base["grads"].items():
            if v is not None:
                scale = float(v.abs().max()) or 1.0
                assert float((a["grads"][k] - v).abs().max()) <= 5e-5 * scale, k     # (float atomics: order-dependent rounding)
        # a second scene with 3x the instances against the stale estimate: must be flagged, not silently truncated
        g2, cam2, bg2 = make_scene(3 * P, W, H, S, 72)
        run_cuda(g2, cam2, bg2)
        with pytest.raises(_C.BinningOverflow):
            _C.check_async(dev, wait=True)
        # the estimate has been raised: the same view now fits and equals the synchronous render
        b = run_cuda(g2, cam2, bg2)
        assert _C.check_async(dev, wait=True) == 1
        _C.set_async_binning(False, dev)
        c = run_cuda(g2, cam2, bg2)
        for k in ("color", "semantics", "depth", "alpha", "radii"):
            assert torch.equal(b[k], c[k]), k
    finally:
        _C.set_async_binning(False, dev)
        _C.device_state(dev).pending.clear()


# ---------------------------------------------------------------------------------------------
# (c) size-independent properties at the full BASELINE config-2 size (1M Gaussians, 1600x1000, S=16)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def big_scene():
    g, cam, bg = make_scene(1_000_000, 1600, 1000, 16, 1)
    return g.to("cuda"), cam.to("cuda"), bg.to("cuda")


def test_full_size_determinism_and_bounds(big_scene):
    g, cam, bg = big_scene
    a = run_cuda(g, cam, bg)
    b = run_cuda(g, cam, bg)
    for k in ("color", "semantics", "depth", "alpha", "radii"):
        assert torch.equal(a[k], b[k]), f"{k} not bit-reproducible"
    assert float(a["alpha"].min()) >= 0.0 and float(a["alpha"].max()) <= 1.0 - 1e-4 + 1e-6
    assert torch.isfinite(a["color"]).all() and torch.isfinite(a["semantics"]).all()


def test_full_size_permutation_invariance(big_scene):
    """Shuffling the Gaussian order must not change the image (random depths => no ties): exercises the
    prefix sum, key emission, radix sort and tile ranges at full size."""
    g, cam, bg = big_scene
    a = run_cuda(g, cam, bg)
    perm = torch.randperm(g.get_xyz.shape[0], device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    gp = SyntheticGaussians(*[t[perm].contiguous() for t in g.tensors()])
    b = run_cuda(gp, cam, bg)
    assert torch.equal(a["radii"][perm], b["radii"])
    # 1M float32 depths in [1,10) do collide now and then; a tie inside one tile composites in index
    # order, which the shuffle changes.  Everything else must agree to rounding.
    for k in ("color", "semantics", "depth", "alpha"):
        d = (a[k] - b[k]).abs().amax(dim=0)
        assert float((d > 1e-5).float().mean()) < 1e-4, k
        assert float(d.max()) < 5e-3, k


def test_full_size_payload_linearity_and_euler(big_scene):
    """The composite is linear in the payload: out(s1 + s2) = out(s1) + out(s2); and for a linear loss
    L = <w, out_sem>,  L = <sem, dL/dsem> (Euler), which ties backward to forward at full size."""
    g, cam, bg = big_scene
    S, H, W = 16, cam.image_height, cam.image_width
    gen = torch.Generator("cuda").manual_seed(17)
    s2 = torch.randn(g.get_semantics.shape, device="cuda", generator=gen)
    o1 = run_cuda(g, cam, bg)
    g2 = SyntheticGaussians(g.get_xyz, g.get_opacity, g.get_scaling, g.get_rotation, g.get_features, s2)
    o2 = run_cuda(g2, cam, bg)
    g12 = SyntheticGaussians(g.get_xyz, g.get_opacity, g.get_scaling, g.get_rotation, g.get_features,
                             g.get_semantics + s2)
    o12 = run_cuda(g12, cam, bg)
    assert float((o12["semantics"] - o1["semantics"] - o2["semantics"]).abs().max()) <= 1e-4

    w = make_loss_weights(S, W, H, 1, device="cuda")
    wz = {k: torch.zeros_like(v) for k, v in w.items()}
    wz["semantics"] = w["semantics"]
    out = run_cuda(g, cam, bg, wz)
    L = float((out["semantics"].double() * w["semantics"].double()).sum())
    euler = float((g.get_semantics.double() * out["grads"]["dL_dsemantics"].double()).sum())
    assert abs(L - euler) <= 1e-3 * max(abs(L), 1.0), (L, euler)


# ---------------------------------------------------------------------------------------------
# gradient accumulation over views (goi_bwd_out.accumulate): the in-place sum the all-reduce operates on
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("use_sh,use_cov,S,P", [(True, False, 16, 5000), (False, True, 5, 5000), (True, False, 0, 5000),
                                                 (True, False, 16, 5003), (True, False, 7, 4999)])
def test_accumulate_mode_sums_views_in_place(use_sh, use_cov, S, P):
    """P % 4 != 0: gradient slots of a flat buffer start at arbitrary float offsets unless padded; GradArena pads
    them, and the second half of the test hands the library deliberately MISALIGNED slots (odd float offsets)."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer, _C
    from goi_b200 import view_parallel as vp
    W, H = 160, 112
    g, cam0, bg = make_scene(P, W, H, S, 21)
    g = g.to("cuda")
    bg = bg.cuda()
    cams = [SyntheticCamera(W, H, math.radians(60.0), _yaw(a), device="cuda") for a in (0.0, 2.0, -3.0)]
    ws = [make_loss_weights(S, W, H, 30 + i, device="cuda") for i in range(3)]
    params = {"means3D": g.get_xyz.clone().requires_grad_(True), "opacities": g.get_opacity.clone().requires_grad_(True)}
    if S:
        params["semantics"] = g.get_semantics.clone().requires_grad_(True)
    if use_sh:
        params["sh"] = g.get_features.clone().requires_grad_(True)
    else:
        params["colors_precomp"] = torch.sigmoid(g.get_features[:, 0, :]).contiguous().requires_grad_(True)
    if use_cov:
        params["cov3D_precomp"] = g.get_covariance(1.0).contiguous().requires_grad_(True)
    else:
        params["scales"] = g.get_scaling.clone().requires_grad_(True)
        params["rotations"] = g.get_rotation.clone().requires_grad_(True)

    def one_view(cam, w):
        settings = GaussianRasterizationSettings(H, W, math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2), bg, 1.0,
                                                 cam.world_view_transform, cam.full_proj_transform, 3,
                                                 cam.camera_center, False, False)
        kw = dict(means3D=params["means3D"], means2D=torch.zeros_like(params["means3D"], requires_grad=True),
                  opacities=params["opacities"], semantics=params.get("semantics"), shs=params.get("sh"),
                  colors_precomp=params.get("colors_precomp"), scales=params.get("scales"),
                  rotations=params.get("rotations"), cov3D_precomp=params.get("cov3D_precomp"))
        color, sem, radii, depth, alpha = GaussianRasterizer(settings)(**kw)
        outs, gr = [color, depth, alpha], [w["render"], w["depth"], w["alpha"]]
        if S:
            outs.append(sem); gr.append(w["semantics"])
        torch.autograd.backward(outs, gr)

    # expected: autograd's own accumulate-add of three independent backward passes
    for cam, w in zip(cams, ws):
        one_view(cam, w)
    expected = {k: p.grad.clone() for k, p in params.items()}
    # in-place: view 0 overwrites the arena, views 1.. add to it
    arena = vp.GradArena(params)
    arena.flat.fill_(float("nan"))               # view 0 must not depend on the previous contents
    for v, (cam, w) in enumerate(zip(cams, ws)):
        arena.clear_grads()
        with arena.accumulating(v > 0):
            one_view(cam, w)
    for k, p in params.items():
        assert p.grad.data_ptr() == arena.slots[k].data_ptr(), f"{k}: .grad does not alias the arena"
        scale = float(expected[k].abs().max()) or 1.0
        err = float((arena.slots[k].view_as(expected[k]) - expected[k]).abs().max())
        assert err <= 1e-4 * scale, f"{k}: accumulated gradient differs by {err / scale:.2e} of max"
        assert arena.slots[k].data_ptr() % 16 == 0, f"{k}: arena slot not 16-byte aligned"
    # the C ABI takes any 4-byte-aligned output pointer: slots packed back to back at ODD float offsets
    flat = torch.full((sum(p.numel() for p in params.values()) + len(params) + 1,), float("nan"), device="cuda")
    slots, off = {}, 1
    for k, p in params.items():
        slots[k] = flat[off:off + p.numel()]
        off += p.numel() + (1 if (off + p.numel()) % 2 == 0 else 0)      # keep every start odd
    assert all(t.data_ptr() % 16 != 0 for t in slots.values())
    for v, (cam, w) in enumerate(zip(cams, ws)):
        for p in params.values():
            p.grad = None
        _C.set_grad_arena(slots, accumulate=v > 0)
        try:
            one_view(cam, w)
        finally:
            _C.set_grad_arena(None)
    for k in params:
        scale = float(expected[k].abs().max()) or 1.0
        err = float((slots[k].view_as(expected[k]) - expected[k]).abs().max())
        assert err <= 1e-4 * scale, f"{k} (misaligned slot): {err / scale:.2e} of max"
