"""A small fused-semantic-loss call (both gt layouts, both precisions, ragged sizes) for compute-sanitizer."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goi-hyperplane_b200"))
import torch  # noqa: E402
from goi_b200.semantic_loss import semantic_loss  # noqa: E402
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1)
for (N, S, K, D) in ((700, 16, 300, 256), (333, 10, 37, 24), (130, 32, 512, 64)):
    x = torch.randn(N, S, generator=g).to(dev).requires_grad_(True)
    W = (torch.randn(K, S, generator=g) * 0.4).to(dev).requires_grad_(True)
    b = (torch.randn(K, generator=g) * 0.1).to(dev).requires_grad_(True)
    lut = (torch.randn(K, D, generator=g) * 0.5 + 0.1).to(dev).requires_grad_(True)
    gt = torch.randn(N, D, generator=g).to(dev)
    for planar in (False, True):
        for prec in (0, 1):
            gg = gt.t().contiguous().reshape(D, 1, N) if planar else gt
            xx = x.t().contiguous().reshape(S, 1, N) if planar else x
            loss, _ = semantic_loss(xx, (W, b), lut, gg, iteration=1, precision=prec)
            loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
