"""Numerical claim behind goi_mask_mma.cuh, checked on the CPU: the 3xTF32 split (x = hi + lo, w = hi + lo;
lo*hi + hi*lo + hi*hi with fp32 accumulation) reproduces an fp32 dot product to ~1e-6 of the sum of |terms|, where a
single TF32 product is only good to ~1e-3 -- so the mask's arg-max over the codebook logits can differ from an fp32
FMA chain only where the top two logits are closer than that (the GPU tests exempt gaps < 1e-4)."""
import numpy as np


def to_tf32(x):
    """cvt.rna.tf32.f32: keep 10 mantissa bits, round to nearest, ties away from zero."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def dot_3xtf32(x, w):
    xh, wh = to_tf32(x), to_tf32(w)
    xl, wl = to_tf32(x - xh), to_tf32(w - wh)
    acc = np.zeros(x.shape[:-1], np.float32)
    for a, b in ((xl, wh), (xh, wl), (xh, wh)):             # small terms first, like the kernel
        acc = acc + np.sum((a.astype(np.float64) * b.astype(np.float64)), axis=-1).astype(np.float32)
    return acc


def test_tf32_rounding_keeps_ten_mantissa_bits():
    x = np.float32([1.0, 1.0 + 2.0 ** -10, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -12, -3.14159265, 1e-30, 65504.0])
    t = to_tf32(x)
    assert np.all((t.view(np.uint32) & 0x1FFF) == 0)
    assert t[1] == x[1] and t[2] == np.float32(1.0 + 2.0 ** -10) and t[3] == np.float32(1.0)   # tie rounds away from zero
    assert np.all(np.abs(t - x) <= np.abs(x) * 2.0 ** -11)


def test_three_term_split_reaches_fp32_accuracy():
    rng = np.random.default_rng(0)
    for S in (4, 10, 16, 32, 64):
        x = rng.standard_normal((4000, S)).astype(np.float32) * rng.uniform(0.01, 30.0, (4000, 1)).astype(np.float32)
        w = (rng.standard_normal((4000, S)) * 0.4).astype(np.float32)
        exact = np.sum(x.astype(np.float64) * w.astype(np.float64), axis=-1)
        scale = np.sum(np.abs(x.astype(np.float64) * w.astype(np.float64)), axis=-1)
        err3 = np.abs(dot_3xtf32(x, w) - exact) / scale
        err1 = np.abs(np.sum(to_tf32(x).astype(np.float64) * to_tf32(w).astype(np.float64), axis=-1) - exact) / scale
        fp32 = np.abs(np.sum(x * w, axis=-1, dtype=np.float32) - exact) / scale
        assert err3.max() <= 2e-6, (S, err3.max())
        assert err1.max() >= 50 * err3.max()                   # a single TF32 product is far worse
        assert err3.max() <= 20 * max(fp32.max(), 1e-7)        # same class as an fp32 accumulation


def test_argmax_agrees_except_on_near_ties():
    rng = np.random.default_rng(1)
    S, K, N = 16, 300, 3000
    x = rng.standard_normal((N, S)).astype(np.float32)
    W = (rng.standard_normal((K, S)) * 0.4).astype(np.float32)
    b = (rng.standard_normal(K) * 0.1).astype(np.float32)
    ref = x.astype(np.float64) @ W.T.astype(np.float64) + b
    got = np.stack([dot_3xtf32(x, np.broadcast_to(W[k], x.shape)) for k in range(K)], axis=1) + b
    top2 = np.sort(ref, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-4
    assert clear.mean() > 0.98
    assert np.array_equal(got.argmax(1)[clear], ref.argmax(1)[clear])
