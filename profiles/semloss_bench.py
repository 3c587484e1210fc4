"""Timing of the fused semantic loss (row f2) next to the reference's torch expression chain on the same GPU.

    python profiles/semloss_bench.py [H W S K D]         (default 1000 1600 16 300 256 = bench config c2)

Prints one JSON line: ms per forward+backward for (a) libgoi_semloss fp32, (b) libgoi_semloss TF32 GEMMs,
(c) the reference chain train.py:142-170 in torch (restated in oracle/semloss_oracle.py, here on the GPU with autograd).
Measurement only; never part of the product path."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goi-hyperplane_b200"))
import torch  # noqa: E402
from torch.nn.functional import cosine_similarity, log_softmax, softmax  # noqa: E402
from goi_b200.semantic_loss import semantic_loss  # noqa: E402

H, W, S, K, D = [int(v) for v in (sys.argv[1:6] if len(sys.argv) >= 6 else (1000, 1600, 16, 300, 256))]
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
N = H * W
sem = torch.randn(S, H, W, generator=g).to(dev).requires_grad_(True)
Wm = (torch.randn(K, S, generator=g) * 0.4).to(dev).requires_grad_(True)
bm = (torch.randn(K, generator=g) * 0.1).to(dev).requires_grad_(True)
lut = (torch.randn(K, D, generator=g) * 0.5 + 0.1).to(dev).requires_grad_(True)
ape = torch.empty(D, H, W, device=dev)
ape.copy_((lut.detach()[torch.randint(0, K, (N,), generator=g).to(dev)] + 0.3 * torch.randn(N, D, device=dev)).t().reshape(D, H, W))


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def ours(precision):
    def f():
        for t in (sem, Wm, bm, lut):
            t.grad = None
        loss, _ = semantic_loss(sem, (Wm, bm), lut, ape, iteration=1, precision=precision)
        loss.backward()
        return loss
    return f


def reference():
    for t in (sem, Wm, bm, lut):
        t.grad = None
    sem_feature = sem.permute(1, 2, 0).reshape(-1, S)
    sem_label = softmax(torch.nn.functional.linear(sem_feature, Wm, bm), dim=-1)
    gtl = ape.float().permute(1, 2, 0).reshape(-1, D)
    gtl = gtl / gtl.norm(dim=1, keepdim=True)
    lut1 = lut / lut.norm(dim=1, keepdim=True)
    sim = gtl @ lut1.T
    sim_val = sim.max(dim=1, keepdim=True)[0]
    label = (sim == sim_val).float().detach()
    lab = torch.nn.MSELoss()(sem_label, label) * 50
    sl = 1 - sim_val.mean()
    recc = 1 - cosine_similarity(lut[sem_label.argmax(-1)], gtl, dim=-1).mean()
    anneal = sim * 1
    b = softmax(anneal, dim=1) * log_softmax(anneal, dim=1)
    sl1 = -1.0 * b.sum(dim=-1).mean()
    loss = lab + sl + 0.3 * sl1 + recc
    loss.backward()
    return loss


res = {"config": dict(H=H, W=W, S=S, K=K, D=D), "unit": "ms per fwd+bwd"}
l32 = float(ours(0)())
res["ours_fp32_ms"] = round(timeit(ours(0)), 3)
ltf = float(ours(1)())
res["ours_tf32_ms"] = round(timeit(ours(1)), 3)
torch.cuda.reset_peak_memory_stats()
lref = float(reference())
res["reference_torch_ms"] = round(timeit(reference, n=3, warm=1), 3)
res["reference_peak_mem_GB"] = round(torch.cuda.max_memory_allocated() / 2**30, 2)
res["loss"] = dict(ours_fp32=l32, ours_tf32=ltf, reference=lref)
print(json.dumps(res))
