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
