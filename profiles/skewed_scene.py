"""Stress: the c2 scene with most Gaussians concentrated in the bottom rows of the image (a ground plane close to
the camera): tile lists are very uneven and the heavy tiles come LAST in row-major block order.  Per-stage times.
    python profiles/skewed_scene.py [fraction_moved] [band]
Measurement only."""
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

frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.6
band = float(sys.argv[2]) if len(sys.argv) > 2 else 0.15
P, W, H, S, seed = CONFIGS["c2"]
dev = torch.device("cuda", 0)
g, _, bg = make_scene(P, W, H, S, seed)
with torch.no_grad():
    n = int(frac * P)
    # camera at the origin looking down +z, y grows downwards on screen: squeeze y/z into the bottom band
    z = g._xyz[:n, 2]
    ymax = z * (H / W) * 0.5773502691896257            # tan(30 deg) * H / W
    t = torch.rand(n, generator=torch.Generator().manual_seed(5))
    g._xyz[:n, 1] = ymax * (1.0 - band * t)
g = g.to(dev).requires_grad_(True)
bg = bg.to(dev)
cam = make_views(W, H, dev)[0]
w = make_loss_weights(S, W, H, seed, device=dev)
outs = ("render", "semantics", "depth", "alpha")


def step():
    for t_ in g.tensors():
        t_.grad = None
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
print(json.dumps({"fraction_moved": frac, "band": band, "num_rendered": _C.num_rendered(),
                  "ms_per_view": round(e0.elapsed_time(e1) / 10, 3),
                  "stages_ms": {k: round(v, 4) for k, v in _C.timing_read().items()}}))
