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
