"""Timing of the hyperplane mask (row a13) and of the fused render+mask (row f4) at bench size.

    python profiles/mask_bench.py [config]        (default c2: 1600x1000, S=16, K=300, D=256)
Prints one JSON line: ms for goi_mask on the rendered [S,H,W] features, for render alone, render followed by goi_mask,
and the fused goi_forward_mask with and without the semantic image; the reference torch chain (gui/main.py:363-385)
beside it.  Measurement only."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goi-hyperplane_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from bench import CONFIGS, make_views  # noqa: E402
from gaussian_renderer import render, render_mask  # noqa: E402
from goi_b200.scenes import PipeFlags, make_mask_model, make_scene  # noqa: E402
from goi_b200.semantic_mask import SemanticHyperplane  # noqa: E402
from common import torch_reference_similarity  # noqa: E402  (tests/common.py)

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
P, W, H, S, seed = CONFIGS[cfg]
dev = torch.device("cuda", 0)
g, _, bg = make_scene(P, W, H, S, seed)
g, bg = g.to(dev), bg.to(dev)
cam = make_views(W, H, dev)[0]
mlp_w, mlp_b, lut, text = (t.to(dev) for t in make_mask_model(S, seed=seed))
hp = SemanticHyperplane(mlp_w, mlp_b, lut, text, thresh=0.86)
pipe = PipeFlags()


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 4)


with torch.no_grad():
    sem = render(cam, g, pipe, bg)["semantics"]
    res = {"config": cfg, "N": W * H, "S": S, "K": int(lut.shape[0]), "unit": "ms"}
    res["mask_kernel"] = timeit(lambda: hp.compute_similarity(sem, channels_first=True))
    res["render"] = timeit(lambda: render(cam, g, pipe, bg))
    res["render_then_mask"] = timeit(lambda: hp.compute_similarity(render(cam, g, pipe, bg)["semantics"], channels_first=True))
    res["fused_render_mask"] = timeit(lambda: render_mask(cam, g, pipe, bg, hp, want_semantics=True))
    res["fused_mask_only"] = timeit(lambda: render_mask(cam, g, pipe, bg, hp, want_semantics=False))
    x = sem.permute(1, 2, 0).reshape(-1, S).contiguous()
    torch.cuda.reset_peak_memory_stats()
    res["reference_torch_chain"] = timeit(lambda: torch_reference_similarity(x, mlp_w, mlp_b, lut, text, thresh=0.86), n=5, warm=1)
    res["reference_peak_mem_GB"] = round(torch.cuda.max_memory_allocated() / 2**30, 2)
print(json.dumps(res))
