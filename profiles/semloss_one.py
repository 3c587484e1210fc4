"""One fused semantic-loss call at bench size (the command ncu wraps for row f2)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goi-hyperplane_b200"))
import torch  # noqa: E402
from goi_b200.semantic_loss import semantic_loss  # noqa: E402
H, W, S, K, D = 1000, 1600, 16, 300, 256
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 0
planar = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
sem = torch.randn(S, H, W, generator=g).to(dev).requires_grad_(True)
Wm = (torch.randn(K, S, generator=g) * 0.4).to(dev).requires_grad_(True)
bm = (torch.randn(K, generator=g) * 0.1).to(dev).requires_grad_(True)
lut = (torch.randn(K, D, generator=g) * 0.5 + 0.1).to(dev).requires_grad_(True)
ape = torch.randn(D, H, W, device=dev) if planar else torch.randn(H * W, D, device=dev)
for _ in range(2):
    loss, _ = semantic_loss(sem, (Wm, bm), lut, ape, iteration=1, precision=prec)
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
