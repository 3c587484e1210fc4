"""bench.py helper: the cpu_baseline leg (the ONE place outside tests/ and smoke() that runs oracle/)."""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "goi-hyperplane_b200"))


def oracle_fwd_bwd(P, W, H, S, seed):
    """Seconds for one forward+backward of the CPU oracle on a make_scene(P, W, H, S, seed) view."""
    from oracle import oracle
    from goi_b200.scenes import make_loss_weights, make_scene
    g, cam, bg = make_scene(P, W, H, S, seed)
    w = {k: v.numpy() for k, v in make_loss_weights(S, W, H, seed).items()}
    kw = dict(means3D=g.get_xyz.numpy(), opacities=g.get_opacity.numpy(), shs=g.get_features.numpy(),
              semantics=g.get_semantics.numpy(), scales=g.get_scaling.numpy(), rotations=g.get_rotation.numpy(),
              W=W, H=H, viewmatrix=cam.world_view_transform.numpy(), projmatrix=cam.full_proj_transform.numpy(),
              campos=cam.camera_center.numpy(), tanfovx=math.tan(cam.FoVx / 2), tanfovy=math.tan(cam.FoVy / 2),
              bg=bg.numpy())
    oracle.lib()
    t0 = time.perf_counter()
    res = oracle.forward(**kw)
    oracle.backward(res, w["render"], w["semantics"], w["depth"], w["alpha"], wide=False)
    t = time.perf_counter() - t0
    res.free()
    # forward tiles run under OpenMP (all cores), the reverse walk is the sequential restatement
    return t, os.cpu_count() or 1
