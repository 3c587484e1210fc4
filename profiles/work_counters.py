"""Work counters of the composite kernels on a bench config (instrumented build, measurement only).

    python goi-hyperplane_b200/build.py --stats
    GOI_RASTER_LIB=goi-hyperplane_b200/lib/libgoi_raster_stats.so python profiles/work_counters.py c2

Prints, for forward and backward: (warp, instance) cull tests, cull survivors (warp walks), walks with at
least one blending lane, and blending (pixel, instance) pairs -- i.e. lane efficiency of the lock-step walk.
"""
import ctypes as C
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

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
P, W, H, S, seed = CONFIGS[cfg]
dev = torch.device("cuda", 0)
L = _C.lib()
g, _, bg = make_scene(P, W, H, S, seed)
g = g.to(dev).requires_grad_(True)
bg = bg.to(dev)
cam = make_views(W, H, dev)[0]
w = make_loss_weights(S, W, H, seed, device=dev)
outs = ("render", "semantics", "depth", "alpha")
buf = (C.c_ulonglong * 8)()
for fn in (L.goi_debug_work_fwd, L.goi_debug_work_bwd):
    fn.restype, fn.argtypes = C.c_int, [C.POINTER(C.c_ulonglong), C.c_int]
    fn(buf, 1)
out = render(cam, g, PipeFlags(), bg)
torch.autograd.backward([out[k] for k in outs], [w[k] for k in outs])
torch.cuda.synchronize()
res = {"config": cfg, "P": P, "W": W, "H": H, "S": S, "num_rendered": _C.num_rendered()}
for name, fn, base in (("fwd", L.goi_debug_work_fwd, 0), ("bwd", L.goi_debug_work_bwd, 4)):
    fn(buf, 0)
    t, walk, anyhit, pairs = (int(buf[base + i]) for i in range(4))
    res[name] = {"cull_tests": t, "warp_walks": walk, "walks_with_hit": anyhit, "blend_pairs": pairs,
                 "lane_efficiency": round(pairs / (32.0 * max(anyhit, 1)), 4),
                 "walks_per_instance": round(walk / max(_C.num_rendered(), 1), 3)}
print(json.dumps(res))
