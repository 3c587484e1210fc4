"""The command compute-sanitizer wraps: one BASELINE config-1-sized forward + backward through the
public API (warp-autonomous composites, tensor-core reductions), an accumulate-mode second view, the asynchronous
forward, and the tcgen05 mask kernel on the rendered features."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "goi-hyperplane_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import common  # noqa: E402
from diff_gaussian_rasterization import _C  # noqa: E402
from goi_b200.scenes import make_loss_weights, make_mask_model, make_scene  # noqa: E402
from goi_b200.semantic_mask import SemanticHyperplane  # noqa: E402

for (P, W, H, S, seed) in ((10_000, 256, 256, 10, 0), (6_000, 250, 197, 16, 2), (3_000, 128, 80, 32, 3)):
    g, cam, bg = make_scene(P, W, H, S, seed)
    w = make_loss_weights(S, W, H, seed)
    out = common.run_cuda(g, cam, bg, w)
    dev = torch.device("cuda", 0)
    _C.set_async_binning(True, dev)
    out2 = common.run_cuda(g, cam, bg, w)
    _C.check_async(dev, wait=True)
    _C.set_async_binning(False, dev)
    d = float((out["color"] - out2["color"]).abs().max())
    print("async vs sync max |d color|", d, "R", _C.num_rendered())
    assert d == 0.0
    mlp_w, mlp_b, lut, text = make_mask_model(S, seed=seed)
    hp = SemanticHyperplane(mlp_w.cuda(), mlp_b.cuda(), lut.cuda(), text.cuda(), thresh=0.86)
    sim = hp.compute_similarity(out["semantics"].detach(), channels_first=True)
    torch.cuda.synchronize()
    print("ok", P, W, H, S, int((sim > 0).sum()))
