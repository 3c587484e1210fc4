"""Golden vectors of the training-side semantic loss, produced by the reference's OWN source lines.

    python tests/golden/make_semloss_golden.py        (authoring container only: needs /root/reference)

The loss lives inline in the reference's training() function (train.py:142-167), so it cannot be imported.  This
script cuts exactly those lines out of /root/reference/train.py (checking a few anchors so a moved block is noticed),
re-targets `.to("cuda")` to the CPU, and executes them with:
    semantic_MLP = the reference's own scene/semantic_model.py:SemanticModel(num_layer=1, use_bias=True) (train.py:64),
    lut, viewpoint_cam.semantic['ape'] ([D,H,W]), sem_feature ([S,H,W]), dataset.sem_dim / ape_dim, iteration
then calls loss.backward() (train.py:170) and stores inputs, loss terms and gradients as tests/golden/semloss_*.npz.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
from torch.nn.functional import cosine_similarity, log_softmax, softmax  # noqa: F401  (names the cut lines use)

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_loss_block():
    lines = open(os.path.join(REF, "train.py")).read().split("\n")
    block = lines[141:163]                      # train.py:142-163 (1-based, inclusive)
    text = "\n".join(block)
    for anchor in ("sem_feature = sem_feature.permute(1, 2, 0).reshape(-1, dataset.sem_dim)",
                   "sim = gtl @ lut1.T", "sem_loss = lab + sl + 0.3 * sl1 + recc"):
        assert anchor in text, f"reference block moved: {anchor!r} not in train.py:142-163"
    import textwrap
    return textwrap.dedent(text).replace('.to("cuda")', '.to("cpu")')


def load_semantic_model():
    spec = importlib.util.spec_from_file_location("ref_semantic_model", os.path.join(REF, "scene", "semantic_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.SemanticModel


def make_case(H, W, S, K, D, iteration, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    SemanticModel = load_semantic_model()
    torch.manual_seed(seed)
    mlp = SemanticModel(dim_in=S, dim_out=K, num_layer=1, use_bias=True, device="cpu").to(dtype)
    with torch.no_grad():
        mlp.layers[0].bias.copy_(torch.randn(K, generator=g).to(dtype) * 0.1)
    lut = torch.nn.Parameter((torch.randn(K, D, generator=g) * 0.5 + 0.1).to(dtype))
    # per-pixel targets: noisy copies of codebook rows (like APE features clustered by k-means, train.py:77-84)
    pick = torch.randint(0, K, (H * W,), generator=g)
    ape = (lut.detach()[pick] + 0.3 * torch.randn(H * W, D, generator=g).to(dtype)) * (0.5 + torch.rand(H * W, 1, generator=g).to(dtype))
    ape = ape.reshape(H, W, D).permute(2, 0, 1).contiguous()
    sem_feature = torch.randn(S, H, W, generator=g).to(dtype).requires_grad_(True)
    cam = types.SimpleNamespace(semantic={"ape": ape})
    ns = dict(torch=torch, softmax=softmax, log_softmax=log_softmax, cosine_similarity=cosine_similarity,
              sem_feature=sem_feature, semantic_MLP=mlp, viewpoint_cam=cam, lut=lut, iteration=iteration,
              dataset=types.SimpleNamespace(sem_dim=S, ape_dim=D))
    exec(reference_loss_block(), ns)
    ns["sem_loss"].backward()
    n = lambda t: t.detach().to(torch.float64).numpy()
    return dict(meta=np.array([H, W, S, K, D, iteration, seed]), sem_feature=n(sem_feature), ape=n(ape),
                mlp_weight=n(mlp.layers[0].weight), mlp_bias=n(mlp.layers[0].bias), lut=n(lut),
                loss=n(ns["sem_loss"]), lab=n(ns["lab"]), sl=n(ns["sl"]), sl1=n(ns["sl1"]), recc=n(ns["recc"]),
                d_sem_feature=n(sem_feature.grad), d_mlp_weight=n(mlp.layers[0].weight.grad),
                d_mlp_bias=n(mlp.layers[0].bias.grad), d_lut=n(lut.grad))


if __name__ == "__main__":
    cases = [("a", 12, 20, 10, 300, 256, 1, 11), ("b", 9, 7, 16, 300, 256, 1500, 12), ("c", 5, 8, 4, 37, 24, 10, 13)]
    for name, H, W, S, K, D, it, seed in cases:
        out = make_case(H, W, S, K, D, it, seed, torch.float32)   # the block itself casts the targets with .float()
        path = os.path.join(HERE, f"semloss_{name}.npz")
        np.savez_compressed(path, **{k: (v.astype(np.float32) if v.dtype == np.float64 and k != "meta" else v) for k, v in out.items()})
        print(path, {k: float(out[k]) for k in ("loss", "lab", "sl", "sl1", "recc")}, os.path.getsize(path) // 1024, "KiB")
