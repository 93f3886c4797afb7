"""CPU-side numerics behind two kernel design decisions (no GPU, no product code: numpy emulation of TF32 rounding).

1. A single TF32 pass (what the tensor-core convolutions use, like the reference's cuDNN default) has ~5e-4 relative error per
   contraction - inside the 1e-3 parity bar, but not "strict fp32" (the reference's torch.bmm in AttnBlock / attention).
2. The operand split of csrc/contract_tc3.cu (a_hi.b_hi + a_lo.b_hi + a_hi.b_lo, fp32 accumulation) recovers fp32-level
   accuracy, which is the premise for moving those strict-fp32 contractions onto the tensor cores."""
import numpy as np


def tf32_rna(x):
    """cvt.rna.tf32.f32: keep 10 mantissa bits, round to nearest, ties away from zero."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))


def test_tf32_single_pass_vs_three_way_split():
    rng = np.random.default_rng(0)
    q = rng.standard_normal((256, 512)).astype(np.float32)
    k = rng.standard_normal((256, 512)).astype(np.float32)
    ref = q.astype(np.float64) @ k.astype(np.float64).T
    fp32 = q @ k.T
    qh, kh = tf32_rna(q), tf32_rna(k)
    ql, kl = tf32_rna(q - qh), tf32_rna(k - kh)
    one = (qh.astype(np.float64) @ kh.astype(np.float64).T).astype(np.float32)
    three = (qh.astype(np.float64) @ kh.astype(np.float64).T + ql.astype(np.float64) @ kh.astype(np.float64).T
             + qh.astype(np.float64) @ kl.astype(np.float64).T).astype(np.float32)
    e1, e3, e32 = _rel(one, ref), _rel(three, ref), _rel(fp32, ref)
    assert 5e-5 < e1 < 1e-3            # single TF32 pass: visible, inside the conv parity bar
    assert e3 < 2e-6 and e3 < 20 * max(e32, 1e-8)   # three-way split: fp32-level
    # the split is exact for the retained part: hi + lo reproduces the input to ~2^-21
    assert float(np.max(np.abs((qh + ql) - q) / (np.abs(q) + 1e-30))) < 2.0 ** -20


def f16_scaled(x, amax=None):
    """The fp16 operand conversion of the convolution kernels (csrc/contract_tc.cu: operand_scale + cvt.rn.satfinite.f16):
    x * s rounded to fp16 (round-to-nearest-even, saturating), s = 2^(14 - floor(log2 amax)); returns the rounded values
    mapped back (x_h / s) so they can be compared with x directly."""
    x = np.asarray(x, dtype=np.float32)
    s = np.float32(1.0)
    if amax is not None and amax > 0 and np.isfinite(amax):
        s = np.float32(2.0) ** np.float32(14 - int(np.floor(np.log2(amax))))
    h = np.clip(x * s, -65504.0, 65504.0).astype(np.float16)
    return h.astype(np.float32) / s


def test_fp16_operands_match_tf32_operand_rounding():
    """fp16 has TF32's 11-bit significand: a contraction with fp16-rounded operands (fp32 accumulate) carries the same
    rounding error as the TF32 pass, for O(1) activations AND - with the power-of-two amax scale - for gradient-sized
    operands (1e-7, far below the fp16 normal range) with a heavy tail."""
    rng = np.random.default_rng(1)
    w = (rng.standard_normal((128, 1152)) * 0.03).astype(np.float32)
    for scale, tail in ((1.0, False), (1.6e-7, True), (3e4, False)):
        a = rng.standard_normal((512, 1152)).astype(np.float32)
        if tail:
            a *= np.exp(2.5 * rng.standard_normal(a.shape)).astype(np.float32)     # log-normal magnitudes: ~8 decades
        a *= np.float32(scale)
        ref = a.astype(np.float64) @ w.astype(np.float64).T
        e_tf32 = _rel((tf32_rna(a).astype(np.float64) @ tf32_rna(w).astype(np.float64).T).astype(np.float32), ref)
        ah = f16_scaled(a, float(np.abs(a).max()))
        wh = f16_scaled(w)
        e_f16 = _rel((ah.astype(np.float64) @ wh.astype(np.float64).T).astype(np.float32), ref)
        assert e_f16 < 6e-4, (scale, e_f16)
        assert e_f16 < 1.5 * e_tf32 + 2e-5, (scale, e_f16, e_tf32)
    # without the scale, gradient-sized operands sit in fp16's subnormal range (why the kernels carry an amax)
    a = (rng.standard_normal((64, 1152)) * 1.6e-7).astype(np.float32)
    assert _rel(f16_scaled(a), a.astype(np.float64)) > 0.05      # 100x the scaled error
    # elements more than 2^-24 below the largest one land in fp16's subnormal range: absolute error <= 2^-25 * 2^-14 of the
    # scaled maximum, i.e. far below the fp32 rounding of the products they take part in
    a = np.array([1.0, 2.0 ** -20, 2.0 ** -26, 3.0], dtype=np.float32)
    assert np.max(np.abs(f16_scaled(a, 3.0) - a)) <= 2.0 ** -11 * 3.0
