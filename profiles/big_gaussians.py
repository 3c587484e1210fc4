"""Stress: the c2 scene plus a few thousand very large Gaussians (many tiles each) -- per-stage times.
    python profiles/big_gaussians.py [n_big] [scale_factor]
Real scenes have a heavy tail of large splats; the per-Gaussian tile loops (count in preprocess, emission) are the
stages that could develop a long tail.  Measurement only."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goi-hyperplane_b200"))
import torch  # noqa: E402
from bench import CONFIGS, make_views  # noqa: E402
from diff_gaussian_rasterization import _C  # noqa: E402
from gaussian_renderer import render  # noqa: E402
from goi_b200.scenes import PipeFlags, make_loss_weights, make_scene  # noqa: E402

n_big = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
factor = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
P, W, H, S, seed = CONFIGS["c2"]
dev = torch.device("cuda", 0)
g, _, bg = make_scene(P, W, H, S, seed)
with torch.no_grad():
    g._scaling[:n_big] *= factor
    g._opacity[:n_big] = g._opacity[:n_big] * 0.05 + 0.01          # faint, so they do not terminate every pixel
g = g.to(dev).requires_grad_(True)
bg = bg.to(dev)
cam = make_views(W, H, dev)[0]
w = make_loss_weights(S, W, H, seed, device=dev)
outs = ("render", "semantics", "depth", "alpha")


def step():
    for t in g.tensors():
        t.grad = None
    out = render(cam, g, PipeFlags(), bg)
    torch.autograd.backward([out[k] for k in outs], [w[k] for k in outs])


for _ in range(3):
    step()
_C.timing_enable(True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
print(json.dumps({"n_big": n_big, "scale_factor": factor, "num_rendered": _C.num_rendered(),
                  "ms_per_view": round(e0.elapsed_time(e1) / 10, 3),
                  "stages_ms": {k: round(v, 4) for k, v in _C.timing_read().items()}}))
