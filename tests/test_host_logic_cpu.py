"""Host-side logic of the token-transformer units, checked WITHOUT a GPU: ops.L.call is replaced by a numpy emulation of the
C-ABI entries' documented semantics (include/mas_b200.h) reading and writing the CPU tensors' memory through the very pointers
and strides the autograd units pass. What is under test is the pointer / stride / batch arithmetic of ops.py (the two-level
batch of the attention contractions, the row pitch of the cross-entropy entry) - the kernels themselves are tested on the GPU."""
import ctypes

import numpy as np
import pytest
import torch


def _addr(v):
    if v is None:
        return 0
    if isinstance(v, torch.Tensor):
        return v.data_ptr()
    if isinstance(v, ctypes.c_void_p):
        return v.value or 0
    raise TypeError(type(v))


def _f32(addr, n):
    return np.ctypeslib.as_array((ctypes.c_float * n).from_address(addr))


def _i64(addr, n):
    return np.ctypeslib.as_array((ctypes.c_int64 * n).from_address(addr))


def _mat(addr, rows, cols, ld, trans):
    """View of op(X) [rows, cols]: stored [rows][ld] (trans False) or [cols][ld] with rows contiguous (trans True)."""
    if not trans:
        flat = _f32(addr, (rows - 1) * ld + cols)
        return np.lib.stride_tricks.as_strided(flat, (rows, cols), (4 * ld, 4))
    flat = _f32(addr, (cols - 1) * ld + rows)
    return np.lib.stride_tricks.as_strided(flat, (rows, cols), (4, 4 * ld))


class Emu:
    """include/mas_b200.h semantics of the entries the transformer units call."""

    def __init__(self):
        self.names = []

    def __call__(self, name, *a):
        self.names.append(name)
        getattr(self, name)(*a)

    def mas_gemm_batched2(self, A, B, C, M, N, K, outer, batch, lda, ldb, ldc, osa, osb, osc, sa, sb, sc, ta, tb, alpha, impl):
        a0, b0, c0 = _addr(A), _addr(B), _addr(C)
        for o in range(outer):
            for i in range(batch):
                Am = _mat(a0 + 4 * (o * osa + i * sa), M, K, lda, bool(ta))          # op(A) [M,K]
                Bm = _mat(b0 + 4 * (o * osb + i * sb), N, K, ldb, not bool(tb))      # op(B)^T as [N,K]: trans_b = 1 is stored [N][K]
                Cm = _mat(c0 + 4 * (o * osc + i * sc), M, N, ldc, False)
                Cm[...] = alpha * (Am.astype(np.float64) @ Bm.astype(np.float64).T)

    def mas_softmax_causal_forward(self, s, p, mats, rows, cols):
        S = _f32(_addr(s), mats * rows * cols).reshape(mats, rows, cols).astype(np.float64)
        out = _f32(_addr(p), mats * rows * cols).reshape(mats, rows, cols)
        mask = np.tril(np.ones((rows, cols)), cols - rows) > 0
        S = np.where(mask, S, -np.inf)
        e = np.exp(S - S.max(-1, keepdims=True))
        out[...] = e / e.sum(-1, keepdims=True)

    def mas_softmax_backward(self, p, dp, ds, rows, cols, scale):
        P = _f32(_addr(p), rows * cols).reshape(rows, cols).astype(np.float64)
        dP = _f32(_addr(dp), rows * cols).reshape(rows, cols).astype(np.float64)
        out = _f32(_addr(ds), rows * cols).reshape(rows, cols)
        out[...] = P * (dP - (dP * P).sum(-1, keepdims=True)) * scale

    def mas_ce_forward(self, logits, ld, target, loss_rows, lse, out, R, V):
        X = _mat(_addr(logits), R, V, ld, False).astype(np.float64)
        t = _i64(_addr(target), R)
        l = np.log(np.exp(X - X.max(-1, keepdims=True)).sum(-1)) + X.max(-1)
        ok = (t >= 0) & (t < V)
        rows = np.where(ok, l - X[np.arange(R), np.clip(t, 0, V - 1)], 0.0)
        _f32(_addr(lse), R)[...] = l
        _f32(_addr(loss_rows), R)[...] = rows
        o = _f32(_addr(out), 2)
        o[0] = rows[ok].mean() if ok.any() else 0.0
        o[1] = ok.sum()

    def mas_ce_backward(self, logits, ld, target, lse, stat, dloss, dlogits, ldd, R, V):
        X = _mat(_addr(logits), R, V, ld, False).astype(np.float64)
        t = _i64(_addr(target), R)
        l = _f32(_addr(lse), R).astype(np.float64)
        cnt = float(_f32(_addr(stat), 2)[1])
        g = float(_f32(_addr(dloss), 1)[0]) / cnt
        D = _mat(_addr(dlogits), R, V, ldd, False)
        ok = (t >= 0) & (t < V)
        p = np.exp(X - l[:, None])
        p[np.arange(R)[ok], t[ok]] -= 1.0
        D[...] = np.where(ok[:, None], p * g, 0.0)


@pytest.fixture
def emu(monkeypatch):
    from mas_b200 import ops
    e = Emu()
    monkeypatch.setattr(ops.L, "call", e)
    monkeypatch.setattr(ops, "_need_cuda", lambda x: None)
    monkeypatch.setattr(ops, "_tc_on", lambda: False)
    return e


@pytest.mark.parametrize("B,S,heads,hd", [(1, 4, 1, 4), (3, 8, 2, 4), (2, 12, 4, 8)])
def test_causal_attention_unit_strides(emu, B, S, heads, hd):
    """CausalAttentionFn (transformer.py:77-103) forward and backward through the emulated entries == torch autograd of the
    textbook formula on the same fused [B,S,3H] activation: pins the two-level batch strides of all six contractions."""
    from mas_b200 import ops
    H = heads * hd
    g = torch.Generator().manual_seed(B * 100 + S)
    qkv = torch.randn(B, S, 3 * H, generator=g)
    w = torch.randn(B, S, H, generator=g)
    x = qkv.clone().requires_grad_(True)
    y = ops.CausalAttentionFn.apply(x, heads)
    (y * w).sum().backward()
    r = qkv.clone().double().requires_grad_(True)
    q, k, v = [t.view(B, S, heads, hd).permute(0, 2, 1, 3) for t in r.split(H, dim=-1)]
    s = (q @ k.transpose(-1, -2)) / hd ** 0.5
    s = s.masked_fill(~torch.tril(torch.ones(S, S, dtype=torch.bool)), float("-inf"))
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B, S, H)
    (ref * w.double()).sum().backward()
    assert emu.names.count("mas_gemm_batched2") == 6          # one launch per contraction, whatever the batch
    assert torch.allclose(y.detach().double(), ref.detach(), atol=1e-5, rtol=1e-5)
    assert torch.allclose(x.grad.double(), r.grad, atol=1e-5, rtol=1e-5)


def test_gemm2_splits_outer_batches_at_the_grid_limit(emu):
    from mas_b200 import ops
    calls = []
    orig = emu.mas_gemm_batched2
    emu.mas_gemm_batched2 = lambda *a: (calls.append(a[6]), orig(*a))
    A = torch.randn(5, 40000, 1, 2)      # 5 outer x 40000 inner 1x2 matrices: 200000 > 65535 matrices per launch
    Bm = torch.randn(5, 40000, 1, 2)
    C = torch.zeros(5, 40000, 1, 1)
    ops.gemm2(A, Bm, C, 1, 1, 2, 5, 40000, 2, 2, 1, 80000, 80000, 40000, 2, 2, 1, tb=True)
    assert calls == [1, 1, 1, 1, 1]
    assert torch.allclose(C.view(-1), (A * Bm).sum(-1).view(-1), atol=1e-6)


@pytest.mark.parametrize("pitch", [17, 24])
def test_cross_entropy_unit(emu, pitch):
    """CrossEntropyFn: row pitch of a sliced logits view, ignored targets, upstream gradient scaling."""
    from mas_b200 import ops
    R, V = 6, 17
    g = torch.Generator().manual_seed(pitch)
    full = torch.randn(2, 3, pitch, generator=g)
    tgt = torch.randint(0, V, (2, 3), generator=g)
    tgt[0, 1] = -100
    x = full[..., :V].clone() if pitch == V else full[..., :V]
    x = x.detach().requires_grad_(True)
    loss = ops.cross_entropy(x, tgt)
    (loss * 0.5).backward()
    r = full[..., :V].double().detach().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(r.reshape(-1, V), tgt.reshape(-1))
    (ref * 0.5).backward()
    assert abs(float(loss) - float(ref)) < 1e-6
    assert torch.allclose(x.grad.double(), r.grad, atol=1e-6)
