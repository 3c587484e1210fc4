"""CPU oracle of the training-side semantic loss -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(goi-hyperplane_b200/) never does.

It restates, operation for operation, the reference's training step /root/reference/train.py:142-167 with plain
torch CPU ops (the reference itself is a torch expression chain, so torch's own CPU kernels + autograd are the
arithmetic being restated; float64 by default for a tight reference, float32 on request):

    :142  sem_feature = sem_feature.permute(1, 2, 0).reshape(-1, sem_dim)        -> caller passes [N,S]
    :143  sem_label = semantic_MLP(sem_feature)          (SemanticModel(num_layer=1, use_bias=True) = one nn.Linear,
                                                          scene/semantic_model.py:31-44, train.py:64)
    :144  sem_label = softmax(sem_label, dim=-1)
    :145-148  gtl = gt.permute(1,2,0).reshape(-1, ape_dim); gtl /= gtl.norm(dim=1, keepdim=True)
    :149  lut1 = lut / lut.norm(dim=1, keepdim=True)
    :150  sim = gtl @ lut1.T
    :152-153  sim_val = sim.max(dim=1, keepdim=True)[0]; label = (sim == sim_val).float().detach()
    :154  lab = MSELoss()(sem_label, label) * 50
    :155  sl = 1 - sim_val.mean()
    :156  recc = 1 - cosine_similarity(lut[sem_label.argmax(-1)], gtl, dim=-1).mean()
    :157-160  t = 1 if iteration < 1000 else 2; anneal = sim * t; b = softmax(anneal) * log_softmax(anneal);
              sl1 = -b.sum(-1).mean()
    :163  sem_loss = lab + sl + 0.3 * sl1 + recc;  :170 loss.backward()

Pinned against the reference's OWN source lines: tests/golden/make_semloss_golden.py extracts exactly those lines from
train.py, executes them on CPU tensors and stores inputs, loss terms and gradients in tests/golden/semloss_*.npz;
tests/test_semloss.py checks this restatement against those fixtures.
"""
from __future__ import annotations

import torch
from torch.nn.functional import cosine_similarity, log_softmax, softmax


def semantic_loss_reference(sem_feature, mlp_weight, mlp_bias, lut, gt, t=1.0, dtype=torch.float64):
    """sem_feature [N,S], mlp_weight [K,S], mlp_bias [K], lut [K,D], gt [N,D] (un-normalised), t = anneal factor.
    Returns dict(loss, lab, sl, sl1, recc, min_sim_val, d_sem_feature, d_mlp_weight, d_mlp_bias, d_lut)."""
    x = sem_feature.detach().to("cpu", dtype).clone().requires_grad_(True)
    W = mlp_weight.detach().to("cpu", dtype).clone().requires_grad_(True)
    b = mlp_bias.detach().to("cpu", dtype).clone().requires_grad_(True)
    L = lut.detach().to("cpu", dtype).clone().requires_grad_(True)
    gtl = gt.detach().to("cpu", dtype).clone()

    sem_label = torch.nn.functional.linear(x, W, b)                       # :143
    sem_label = softmax(sem_label, dim=-1)                                # :144
    gtl = gtl / gtl.norm(dim=1, keepdim=True)                             # :148
    lut1 = L / L.norm(dim=1, keepdim=True)                                # :149
    sim = gtl @ lut1.T                                                    # :150
    sim_val = sim.max(dim=1, keepdim=True)[0]                             # :152
    label = (sim == sim_val).to(dtype).detach()                           # :153
    lab = torch.nn.MSELoss()(sem_label, label) * 50                       # :154
    sl = 1 - sim_val.mean()                                               # :155
    recc = 1 - cosine_similarity(L[sem_label.argmax(-1)], gtl, dim=-1).mean()   # :156
    anneal = sim * t                                                      # :158
    bb = softmax(anneal, dim=1) * log_softmax(anneal, dim=1)              # :159
    sl1 = -1.0 * bb.sum(dim=-1).mean()                                    # :160
    sem_loss = lab + sl + 0.3 * sl1 + recc                                # :163
    sem_loss.backward()                                                   # :170
    return dict(loss=sem_loss.detach(), lab=lab.detach(), sl=sl.detach(), sl1=sl1.detach(), recc=recc.detach(),
                min_sim_val=sim_val.min().detach(), d_sem_feature=x.grad, d_mlp_weight=W.grad, d_mlp_bias=b.grad,
                d_lut=L.grad)
