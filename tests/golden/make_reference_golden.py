"""Generate golden vectors FROM THE REFERENCE'S OWN CUDA KERNELS (oracle/_ref/libref_S*.so, i.e. the
reference's cuda_rasterizer/*.cu compiled unmodified for sm_100a by oracle/build.py) on a B200.

    gpurun -- python tests/golden/make_reference_golden.py      # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/

Each file holds the seeded scene parameters' identifiers (the scene itself is regenerated from the
seed by goi_b200.scenes.make_scene), the reference's forward outputs and, for the linear pseudo-loss
of make_loss_weights, all of its backward outputs.  tests/test_oracle_golden.py (CPU) checks the C
oracle against these; tests/test_gpu_parity.py checks the CUDA path against them.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goi-hyperplane_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [
    # name, P, W, H, S, seed, use_sh, use_cov
    ("ref_c1small_S10", 600, 96, 64, 10, 101, True, False),
    ("ref_ragged_S16", 500, 83, 61, 16, 102, True, False),
    ("ref_precomp_S1", 400, 64, 48, 1, 103, False, True),
    ("ref_wide_S32", 300, 64, 48, 32, 104, True, False),
]


def main():
    import common
    from goi_b200.scenes import make_loss_weights, make_scene
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, P, W, H, S, seed, use_sh, use_cov in CASES:
        g, cam, bg = make_scene(P, W, H, S, seed)
        bg = torch.tensor([0.25, 0.5, 0.75])
        w = make_loss_weights(S, W, H, seed)
        ref = common.run_reference_cuda(g, cam, bg, w, use_sh=use_sh, use_cov=use_cov)
        arrays = {"meta": np.array([P, W, H, S, seed, int(use_sh), int(use_cov)], dtype=np.int64),
                  "bg": bg.numpy(), "num_rendered": np.array([ref["num_rendered"]], dtype=np.int64)}
        for k in ("color", "semantics", "depth", "alpha", "radii"):
            arrays[k] = ref[k].cpu().numpy()
        for k, v in ref["grads"].items():
            arrays[k] = v.cpu().numpy()
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(name, "R =", ref["num_rendered"], "visible =", int((ref["radii"] > 0).sum()), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
