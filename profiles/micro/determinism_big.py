import os, sys, collections
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "goi-hyperplane_b200"))
import torch
from bench import CONFIGS, make_views
from diff_gaussian_rasterization import _C
from gaussian_renderer import render
from goi_b200.scenes import PipeFlags, make_scene
P, W, H, S, seed = CONFIGS["c2"]
dev = torch.device("cuda", 0)
g, _, bg = make_scene(P, W, H, S, seed)
with torch.no_grad():
    g._scaling[:2000] *= 60.0
    g._opacity[:2000] = g._opacity[:2000] * 0.05 + 0.01
g = g.to(dev); bg = bg.to(dev)
cam = make_views(W, H, dev)[0]
c = collections.Counter(); ref = None; nd = 0
with torch.no_grad():
    for i in range(60):
        out = render(cam, g, PipeFlags(), bg)
        c[_C.num_rendered()] += 1
        if ref is None: ref = out["render"].clone()
        else: nd += int(not torch.equal(ref, out["render"]))
print(dict(c), "image mismatches", nd)
