"""The operand split of the 3xTF32 mode (csrc/precise_tc.cuh:tf32_split, tf32_presplit_kernel) restated in numpy with the
same integer arithmetic -- (bits + 0x1000) & 0xFFFFE000 = round to nearest, ties away, on the magnitude bits -- and the
properties the kernels rely on: both parts are TF32 numbers (13 low mantissa bits clear), hi + lo reproduces x to 2^-22
relative, and the three-product sum hi*hi + lo*hi + hi*lo reproduces the fp32 product to ~2^-21."""
import numpy as np


def _round_tf32(x):
    b = np.asarray(x, np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def _split(x):
    x = np.asarray(x, np.float32)
    hi = _round_tf32(x)
    lo = _round_tf32((x - hi).astype(np.float32))
    return hi, lo


def test_split_parts_are_tf32_and_sum_to_x():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(200000), rng.standard_normal(1000) * 1e-20, rng.standard_normal(1000) * 1e20,
                        np.array([0.0, -0.0, 1.0, -1.0, 255.0, 2.0 ** -120, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -12])]).astype(np.float32)
    hi, lo = _split(x)
    assert ((hi.view(np.uint32) | lo.view(np.uint32)) & np.uint32(0x1FFF)).max() == 0       # both are TF32 values
    assert np.abs(hi.astype(np.float64) - x).max() <= 2.0 ** -11 * np.abs(x).max()
    err = np.abs((hi.astype(np.float64) + lo.astype(np.float64)) - x.astype(np.float64))
    assert (err <= 2.0 ** -22 * np.abs(x.astype(np.float64)) + 1e-45).all()
    # ties on the dropped bits round away from zero, like cvt.rna.tf32.f32
    tie = np.float32(1.0 + 2.0 ** -11)
    assert _round_tf32(tie) == np.float32(1.0 + 2.0 ** -10) and _round_tf32(-tie) == np.float32(-(1.0 + 2.0 ** -10))


def test_three_products_reproduce_the_fp32_product():
    rng = np.random.default_rng(1)
    a = rng.standard_normal(100000).astype(np.float32)
    w = rng.standard_normal(100000).astype(np.float32)
    ah, al = _split(a)
    wh, wl = _split(w)
    f = lambda t: t.astype(np.float64)
    three = f(ah) * f(wh) + f(al) * f(wh) + f(ah) * f(wl)
    exact = f(a) * f(w)
    rel = np.abs(three - exact) / np.maximum(np.abs(exact), 1e-300)
    assert rel.max() <= 2.0 ** -20 and rel.mean() <= 2.0 ** -23            # dropped al*wl ~ 2^-22, lo roundings 2^-22 each
    one = f(ah) * f(wh)
    assert (np.abs(one - exact) / np.maximum(np.abs(exact), 1e-300)).mean() >= 2.0 ** -13   # a single TF32 product is not enough


def test_truncating_accumulate_model_explains_the_measured_bias():
    """What was measured on the B200 with ONE TMEM accumulator per conv (profiles/r2s5_tf32x3_accuracy_single_accumulator.txt:
    positive operands, result low by ~1.0e-8 x K relative) against a numpy model of the accumulate: the 8 products of an
    MMA instruction summed exactly, the sum added to the fp32 accumulator and TRUNCATED toward zero.  The model predicts
    -0.6e-8 ... -0.8e-8 x K; the hardware loses a little more (it also truncates inside the product tree), a
    round-to-nearest accumulate would lose ~100x less.  This is why the kernels fold chunks of <= 36 instructions into
    registers with round-to-nearest adds."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "profiles", "r2s5_tf32x3_accuracy_single_accumulator.txt")).read()
    pos = txt.split("inputs signed")[0]
    measured = {int(m.group(1)): float(m.group(2)) for m in
                re.finditer(r"1x1 Cin\s+\d+\s+K\s+(\d+)\s+fp32:.*?tf32x3: max \S+ bias (\S+)", pos)}
    assert set(measured) >= {128, 512, 2048}

    def trunc32(x64):
        f = x64.astype(np.float32)
        over = np.abs(f.astype(np.float64)) > np.abs(x64)
        return np.where(over, np.nextafter(f, np.float32(0)), f).astype(np.float32)

    rng = np.random.default_rng(0)
    for K in (128, 512, 2048):
        n = 600
        a = np.abs(rng.standard_normal((n, K))).astype(np.float32)
        w = np.abs(rng.standard_normal((n, K)) * np.sqrt(2.0 / K)).astype(np.float32)
        (ah, al), (wh, wl) = _split(a), _split(w)
        f = lambda t: t.astype(np.float64)
        exact = (f(a) * f(w)).sum(1)
        acc_t = np.zeros(n, np.float32)
        acc_n = np.zeros(n, np.float32)
        for k0 in range(0, K, 8):
            s = slice(k0, k0 + 8)
            for x, y in ((ah, wh), (al, wh), (ah, wl)):
                p = (f(x[:, s]) * f(y[:, s])).sum(1)
                acc_t = trunc32(acc_t.astype(np.float64) + p)
                acc_n = (acc_n.astype(np.float64) + p).astype(np.float32)
        bias_t = ((acc_t.astype(np.float64) - exact) / np.abs(exact).max()).mean()
        bias_n = ((acc_n.astype(np.float64) - exact) / np.abs(exact).max()).mean()
        assert bias_t < 0 and 1.0 <= measured[K] / bias_t <= 1.7, (K, measured[K], bias_t)
        assert abs(bias_n) < abs(bias_t) / 20
