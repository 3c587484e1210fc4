"""Diagnostic print-out of the fused semantic loss against the CPU oracle (never asserts): every loss term and gradient,
row-major and planar inputs, both precisions.  python profiles/semloss_diag.py [N S K D]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goi-hyperplane_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from oracle.semloss_oracle import semantic_loss_reference  # noqa: E402
from test_semloss import run_cuda, rel, TERMS, GRADS  # noqa: E402

cfgs = [tuple(int(v) for v in sys.argv[1:5])] if len(sys.argv) >= 5 else [(4099, 16, 300, 256), (777, 32, 512, 64), (33, 3, 5, 8), (1, 16, 300, 256)]
for N, S, K, D in cfgs:
    g = torch.Generator().manual_seed(N + S)
    x = torch.randn(N, S, generator=g)
    W, b = torch.randn(K, S, generator=g) * 0.4, torch.randn(K, generator=g) * 0.1
    lut = torch.randn(K, D, generator=g) * 0.5 + 0.1
    gt = lut[torch.randint(0, K, (N,), generator=g)] + 0.3 * torch.randn(N, D, generator=g)
    ref = semantic_loss_reference(x, W, b, lut, gt, t=1.0, dtype=torch.float64)
    for planar in (None, (1, N)):
        for prec in (0, 1):
            try:
                cu = run_cuda(x, W, b, lut, gt, 1, planar, precision=prec)
            except Exception as e:  # noqa: BLE001
                print(N, S, K, D, "planar" if planar else "rowmajor", prec, "FAILED", repr(e)[:300])
                continue
            line = {k: f"{abs(float(cu[k]) - float(ref[k])):.2e}" for k in TERMS}
            line.update({k: f"{rel(cu[k].numpy(), ref[k].numpy().reshape(cu[k].shape)):.2e}" for k in GRADS})
            print(N, S, K, D, "planar" if planar else "rowmajor", "tf32" if prec else "fp32", line, flush=True)
