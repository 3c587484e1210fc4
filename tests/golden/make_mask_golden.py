"""Golden vectors of the open-vocabulary mask path, produced by the reference's OWN source.

    python tests/golden/make_mask_golden.py        (authoring container only: needs /root/reference)

What runs is the reference's code, unmodified, on the CPU:
  * ``GUI.compute_similarity``         cut out of /root/reference/gui/main.py:363-385 and exec'd (the GUI class itself
                                       cannot be imported: dearpygui / kiui / CLIP are not in this image);
  * ``ApeSimMeasure.compute_similarity``  cut out of gui/main.py:113-117;
  * ``SemanticModel``                  imported by path from scene/semantic_model.py (the 1-layer "semantic_MLP" of
                                       train.py:64, `num_layer=1, use_bias=True`);
  * ``VisionLanguageAlign``            imported by path from ext/vision_language_align.py (its
                                       compute_dot_product_logit_betweenTandI_manualbias, :109-122);
  * ``LinearSVM``                      imported by path from networks.py (the OSH head, :12-59; its `utils.image_utils`
                                       import is stubbed -- only `forward` and the constructor run).
The only edits to the cut text are `.cuda()` -> `.cpu()` (there is no GPU in the authoring container).  The codebook row
each element picked (`sem_logit`, a local of the reference function) and the MLP logits are captured by wrapping
`renderer.LUT` / `renderer.MLP` in recording proxies.  Outputs: tests/golden/mask_*.npz.
"""
import importlib.util
import os
import sys
import textwrap
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _cut(first, last, anchors):
    """Lines first..last (1-based, inclusive) of the reference's gui/main.py, dedented; anchors guard against drift."""
    lines = open(os.path.join(REF, "gui", "main.py")).read().split("\n")
    text = textwrap.dedent("\n".join(lines[first - 1:last]))
    for a in anchors:
        assert a in text, f"reference block moved: {a!r} not in gui/main.py:{first}-{last}"
    return text.replace(".cuda()", ".cpu()")


def reference_functions():
    ns = dict(torch=torch)
    exec(_cut(363, 385, ["def compute_similarity(self, embedding_feature, out_bg_mask=None):",
                         "sem_logit = torch.softmax(dec_feature * 10, dim=-1).argmax(dim=-1)",
                         "sim[_bg_mask] = 0"]), ns)
    gui_compute_similarity = ns["compute_similarity"]
    ns2 = dict(torch=torch)
    exec(_cut(113, 117, ["def compute_similarity(self, semantic_feature):",
                         "compute_dot_product_logit_betweenTandI_manualbias"]), ns2)
    return gui_compute_similarity, ns2["compute_similarity"]


class _RecordingLUT:
    """renderer.LUT stand-in: indexing returns the codebook rows and remembers the index tensor."""
    def __init__(self, lut):
        self.lut, self.last_index = lut, None

    def __getitem__(self, idx):
        self.last_index = idx.clone()
        return self.lut[idx]


class _RecordingMLP:
    def __init__(self, mlp):
        self.mlp, self.last_out = mlp, None

    def __call__(self, x):
        self.last_out = self.mlp(x)
        return self.last_out


def make_case(N, S, K, D, mode, log_scale, thresh, seed):
    SemanticModel = _load("ref_semantic_model", "scene/semantic_model.py").SemanticModel
    VisionLanguageAlign = _load("ref_vla", "ext/vision_language_align.py").VisionLanguageAlign
    img_utils = types.ModuleType("utils.image_utils")
    img_utils.apply_mask = img_utils.compute_mask_ratio = img_utils.calculate_iou = lambda *a, **k: None
    sys.modules.setdefault("utils", types.ModuleType("utils"))
    sys.modules["utils.image_utils"] = img_utils
    LinearSVM = _load("ref_networks", "networks.py").LinearSVM
    gui_cs, ape_cs = reference_functions()

    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    mlp = SemanticModel(dim_in=S, dim_out=K, num_layer=1, use_bias=True, device="cpu")      # train.py:64
    with torch.no_grad():
        mlp.layers[0].bias.copy_((torch.rand(K, generator=g) * 2 - 1) / S ** 0.5)
    lut = torch.randn(K, D, generator=g)
    lut = lut / lut.norm(dim=1, keepdim=True) * (0.5 + torch.rand(K, 1, generator=g))       # rows are NOT unit length
    text = torch.randn(1, D, generator=g)
    text = text / text.norm() * 3.0             # so that sigmoid(f.w / e^ls + 2) straddles the 0.86 threshold
    x = torch.randn(N, S, generator=g)
    x[: N // 8] *= 0.0                          # background pixels render a zero feature vector: logits = biases
    x[N // 8: N // 4] *= 8.0                    # large activations

    cls_embd = VisionLanguageAlign(D, 1024, log_scale=log_scale)
    vlm = types.SimpleNamespace(cls_embd=cls_embd, text_feature=text, device=torch.device("cpu"), feature_dim=D)
    vlm.compute_similarity = types.MethodType(ape_cs, vlm)
    rlut, rmlp = _RecordingLUT(lut), _RecordingMLP(mlp)
    gui = types.SimpleNamespace(renderer=types.SimpleNamespace(MLP=rmlp, LUT=rlut), vlm=vlm, res_finetuned=False,
                                resMLP=None, clip_feature_thresh=thresh)
    svm_w, svm_b = np.zeros((1, D), np.float32), np.zeros((1,), np.float32)
    if mode == "osh":
        # gui/main.py:1676-1680: LinearSVM(set_bias=clip_feature_thresh, input_dim=256), weights = the text feature
        gui.res_finetuned = True
        gui.resMLP = LinearSVM(set_bias=thresh, input_dim=D)
        gui.resMLP.weight_set(text)
        svm_w = gui.resMLP.linear.weight.detach().numpy().copy()
        svm_b = gui.resMLP.linear.bias.detach().numpy().copy()
    bg = torch.zeros(N, dtype=torch.bool)
    sim = gui_cs(gui, x, bg)                    # <- the reference's GUI.compute_similarity
    logits = rmlp.last_out.detach()
    top2 = torch.topk(logits, 2, dim=-1).values
    return dict(meta=np.array([N, S, K, D, 1 if mode == "osh" else 0, seed]), x=x.numpy(),
                mlp_weight=mlp.layers[0].weight.detach().numpy(), mlp_bias=mlp.layers[0].bias.detach().numpy(),
                lut=lut.numpy(), text=text.numpy(), log_scale=np.float32(log_scale), thresh=np.float32(thresh),
                svm_weight=svm_w, svm_bias=svm_b, sim=sim.numpy(), bg_mask=bg.numpy(),
                idx=rlut.last_index.numpy().astype(np.int32), top2_gap=(top2[:, 0] - top2[:, 1]).numpy())


if __name__ == "__main__":
    cases = [("ape_S16", 6000, 16, 300, 256, "ape", 0.2, 0.86, 21), ("osh_S10", 4000, 10, 300, 256, "osh", 0.0, 0.86, 22),
             ("ape_S32", 3000, 32, 300, 256, "ape", 0.0, 0.86, 23), ("osh_S4_small", 1500, 4, 37, 24, "osh", 0.0, 0.7, 24),
             ("ape_S64", 1000, 64, 300, 256, "ape", -0.3, 0.9, 25)]
    for name, N, S, K, D, mode, ls, th, seed in cases:
        out = make_case(N, S, K, D, mode, ls, th, seed)
        path = os.path.join(HERE, f"mask_{name}.npz")
        np.savez_compressed(path, **out)
        print(path, "kept", int((out["sim"] > 0).sum()), "of", N, "| distinct rows", len(np.unique(out["idx"])),
              "| min top-2 gap %.2e" % out["top2_gap"].min(), os.path.getsize(path) // 1024, "KiB")
