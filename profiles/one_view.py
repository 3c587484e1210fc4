"""One forward + backward of a bench config through the public API (the command ncu wraps).

    ncu --set full --clock-control none --import-source on -k regex:k_composite -c 2 -o gpurun_out/prof \
        python profiles/one_view.py c2 [n_views]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goi-hyperplane_b200"))
import torch  # noqa: E402
from bench import CONFIGS, make_views  # noqa: E402
from gaussian_renderer import render  # noqa: E402
from goi_b200.scenes import PipeFlags, make_loss_weights, make_scene  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
P, W, H, S, seed = CONFIGS[cfg]
dev = torch.device("cuda", 0)
g, _, bg = make_scene(P, W, H, S, seed)
g = g.to(dev).requires_grad_(True)
bg = bg.to(dev)
cams = make_views(W, H, dev)
w = make_loss_weights(S, W, H, seed, device=dev)
outs = ("render", "semantics", "depth", "alpha")
for i in range(n):
    out = render(cams[i % len(cams)], g, PipeFlags(), bg)
    torch.autograd.backward([out[k] for k in outs], [w[k] for k in outs])
torch.cuda.synchronize()
print("ok", cfg, n)
