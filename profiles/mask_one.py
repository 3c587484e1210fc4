"""One hyperplane-mask call at bench size on random features (the command ncu wraps for row a13)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goi-hyperplane_b200"))
import torch  # noqa: E402
from goi_b200.scenes import make_mask_model  # noqa: E402
from goi_b200.semantic_mask import SemanticHyperplane  # noqa: E402
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, W = 1000, 1600
dev = torch.device("cuda", 0)
x = torch.randn(S, H, W, device=dev)
hp = SemanticHyperplane(*(t.to(dev) for t in make_mask_model(S, seed=1)), thresh=0.86)
for _ in range(3):
    sim = hp.compute_similarity(x, channels_first=True)
torch.cuda.synchronize()
print("ok", int((sim > 0).sum()))
