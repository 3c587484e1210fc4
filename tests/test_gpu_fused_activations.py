"""f1 row: render(..., fused_activations=True) takes the STORED parameters and must equal
the reference call form -- torch sigmoid / exp / normalize / cat in front of the rasterizer and autograd
through them behind it (scene/gaussian_model.py:90-117, gaussian_renderer/__init__.py:53-95)."""
import math

import pytest
import torch

from gaussian_renderer import render
from goi_b200.gaussian_cloud import GaussianCloud
from goi_b200.scenes import PipeFlags, SyntheticCamera, make_loss_weights, make_scene

pytestmark = pytest.mark.gpu
OUTS = ("render", "semantics", "depth", "alpha")


def _cloud(P, W, H, S, seed):
    g, cam, bg = make_scene(P, W, H, S, seed)
    gain = 0.25 + 3.0 * torch.rand(P, generator=torch.Generator().manual_seed(seed))   # |q| != 1: normalize matters
    c = GaussianCloud.from_activated(g.get_xyz, g.get_opacity, g.get_scaling, g.get_rotation, g.get_features,
                                     g.get_semantics, rotation_gain=gain)
    if S == 0:
        c._semantics = torch.zeros(P, 0)
    return c.to("cuda").requires_grad_(True), cam.to("cuda"), torch.tensor([0.1, 0.3, 0.2], device="cuda")


def _run(c, cam, bg, w, fused, S):
    for t in c.parameters().values():
        t.grad = None
    out = render(cam, c, PipeFlags(), bg, fused_activations=fused)
    keys = [k for k in OUTS if not (k == "semantics" and S == 0)]
    torch.autograd.backward([out[k] for k in keys], [w[k] for k in keys])
    grads = {k: t.grad.clone() for k, t in c.parameters().items() if t.grad is not None}
    grads["means2D"] = out["viewspace_points"].grad.clone()
    return {k: out[k].detach().clone() for k in OUTS + ("radii",)}, grads


@pytest.mark.parametrize("P,W,H,S,seed", [(20_000, 320, 200, 16, 1), (5_000, 250, 131, 10, 2), (3_000, 128, 96, 0, 3),
                                          (4_033, 160, 112, 7, 4)])
def test_fused_equals_reference_call_form(P, W, H, S, seed):
    c, cam, bg = _cloud(P, W, H, S, seed)
    w = make_loss_weights(S, W, H, seed, device="cuda")
    ref_out, ref_g = _run(c, cam, bg, w, False, S)
    fus_out, fus_g = _run(c, cam, bg, w, True, S)
    # radii: exp / normalize differ from torch's kernels by at most an ulp; ceil() may flip on a handful
    assert float((ref_out["radii"] != fus_out["radii"]).float().mean()) < 1e-4
    for k in OUTS:
        if ref_out[k].numel() == 0:
            continue
        d = (ref_out[k] - fus_out[k]).abs().amax(dim=0)
        assert float((d > 1e-4).float().mean()) < 2e-4, f"{k}: {float(d.max()):.3e}"
    assert set(fus_g) == set(ref_g)
    for k in ref_g:
        scale = float(ref_g[k].abs().max()) or 1.0
        err = float((ref_g[k] - fus_g[k]).abs().max())
        assert err <= 1e-3 * scale, f"dL/d{k}: {err / scale:.3e} of max"


@pytest.mark.parametrize("P,W,H,S,seed", [(10_000, 256, 256, 16, 11), (4_033, 160, 112, 7, 12)])
def test_fused_matches_oracle_on_activated_parameters(P, W, H, S, seed):
    """Independent check of the fused path: the CPU oracle renders the ACTIVATED parameters (torch CPU sigmoid / exp /
    normalize / cat of the stored ones), and its gradients are pulled back to the stored parameters through those
    same torch CPU activations -- no CUDA code on the comparison side."""
    import numpy as np
    import common
    c, cam, bg = _cloud(P, W, H, S, seed)
    w = make_loss_weights(S, W, H, seed, device="cuda")
    fus_out, fus_g = _run(c, cam, bg, w, True, S)

    raw = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in c.parameters().items()}
    act = dict(xyz=raw["xyz"], opacity=torch.sigmoid(raw["opacity"]), scaling=torch.exp(raw["scaling"]),
               rotation=torch.nn.functional.normalize(raw["rotation"]),
               features=torch.cat((raw["f_dc"], raw["f_rest"]), dim=1), semantics=raw["semantics"])
    from goi_b200.scenes import SyntheticGaussians
    ga = SyntheticGaussians(*[act[k].detach().contiguous() for k in ("xyz", "opacity", "scaling", "rotation", "features",
                                                                     "semantics")])
    ora = common.run_oracle(ga, cam.to("cpu"), bg.cpu(), {k: v.cpu() for k, v in w.items()})
    # ceil(3 sqrt(lambda)) may flip where exp / normalize differ by an ulp between torch-CPU and the fused kernel
    assert (common.to_np(fus_out["radii"]) != ora["radii"]).mean() < 1e-4
    got = dict(color=fus_out["render"], semantics=fus_out["semantics"], depth=fus_out["depth"], alpha=fus_out["alpha"])
    common.assert_images_close(got, ora, max_bad_frac=3e-4, what="fused activations vs oracle")
    og = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in ora["grads"].items()}
    torch.autograd.backward(
        [act["xyz"], act["opacity"], act["scaling"], act["rotation"], act["features"], act["semantics"]],
        [og["dL_dmeans3D"], og["dL_dopacity"].view_as(act["opacity"]), og["dL_dscales"], og["dL_drotations"],
         og["dL_dsh"].view_as(act["features"]), og["dL_dsemantics"]])
    for k, t in raw.items():
        ref = t.grad
        scale = float(ref.abs().max()) or 1.0
        err = float((fus_g[k].cpu() - ref).abs().max())
        assert err <= 2e-3 * scale, f"dL/d{k}: {err / scale:.3e} of max vs oracle"
    scale = float(np.abs(ora["grads"]["dL_dmeans2D"]).max())
    assert float((fus_g["means2D"].cpu() - og["dL_dmeans2D"]).abs().max()) <= 2e-3 * scale


def test_fused_rejects_non_default_pipeline():
    c, cam, bg = _cloud(500, 64, 48, 4, 5)
    with pytest.raises(ValueError):
        render(cam, c, PipeFlags(compute_cov3D_python=True), bg, fused_activations=True)
    c.set_semantic_masks(torch.ones(500, device="cuda"))
    with pytest.raises(ValueError):
        render(cam, c, PipeFlags(), bg, fused_activations=True)


def test_fused_accumulates_views_in_place():
    """raw opacities go through a per-view scratch (sigmoid' is applied once per view, not to the running sum)."""
    from goi_b200 import view_parallel as vp
    P, W, H, S = 6000, 192, 128, 8
    c, cam0, bg = _cloud(P, W, H, S, 7)
    cams = [cam0, SyntheticCamera(W, H, math.radians(60.0), torch.tensor(
        [[math.cos(0.05), 0, math.sin(0.05), 0], [0, 1, 0, 0], [-math.sin(0.05), 0, math.cos(0.05), 0], [0, 0, 0, 1.0]]),
        device="cuda")]
    ws = [make_loss_weights(S, W, H, 40 + i, device="cuda") for i in range(2)]
    for t in c.parameters().values():
        t.grad = None
    for cam, w in zip(cams, ws):
        out = render(cam, c, PipeFlags(), bg, fused_activations=True)
        torch.autograd.backward([out[k] for k in OUTS], [w[k] for k in OUTS])
    expected = {k: t.grad.clone() for k, t in c.parameters().items()}
    arena = vp.GradArena({"means3D": c._xyz, "opacities": c._opacity, "semantics": c._semantics, "sh": c._features_dc,
                          "sh_rest": c._features_rest, "scales": c._scaling, "rotations": c._rotation})
    arena.flat.fill_(float("nan"))
    for v, (cam, w) in enumerate(zip(cams, ws)):
        arena.clear_grads()
        out = render(cam, c, PipeFlags(), bg, fused_activations=True)
        with arena.accumulating(v > 0):
            torch.autograd.backward([out[k] for k in OUTS], [w[k] for k in OUTS])
    names = {"means3D": "xyz", "opacities": "opacity", "semantics": "semantics", "sh": "f_dc", "sh_rest": "f_rest",
             "scales": "scaling", "rotations": "rotation"}
    for slot, pname in names.items():
        e = expected[pname]
        err = float((arena.slots[slot].view_as(e) - e).abs().max())
        assert err <= 1e-4 * (float(e.abs().max()) or 1.0), f"{slot}: {err:.3e}"
