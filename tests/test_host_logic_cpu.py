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

    def mas_gemm_batched2(self, A, B, C, M, N, K, outer, batch, lda, ldb, ldc, osa, osb, osc, sa, sb, sc, ta, tb, alpha, impl, causal=0):
        a0, b0, c0 = _addr(A), _addr(B), _addr(C)
        for o in range(outer):
            for i in range(batch):
                Am = _mat(a0 + 4 * (o * osa + i * sa), M, K, lda, bool(ta))          # op(A) [M,K]
                Bm = _mat(b0 + 4 * (o * osb + i * sb), N, K, ldb, not bool(tb))      # op(B)^T as [N,K]: trans_b = 1 is stored [N][K]
                Cm = _mat(c0 + 4 * (o * osc + i * sc), M, N, ldc, False)
                full = alpha * (Am.astype(np.float64) @ Bm.astype(np.float64).T)
                if causal in (1, 2):       # the kernel skips blocks the hint declares zero: the operand really has to be zero there
                    Az = Am.astype(np.float64)
                    assert np.all(np.triu(Az, 1) == 0) if causal == 1 else np.all(np.tril(Az, -1) == 0), "causal hint on a non-triangular operand"
                if causal == 3:            # tiles above the diagonal may be left as garbage / zeros: poison them to prove nobody reads them
                    full = np.where(np.tril(np.ones((M, N))) > 0, full, np.nan)
                Cm[...] = full

    def mas_softmax_causal_forward(self, s, p, mats, rows, cols):
        S = _f32(_addr(s), mats * rows * cols).reshape(mats, rows, cols).astype(np.float64)
        out = _f32(_addr(p), mats * rows * cols).reshape(mats, rows, cols)
        mask = np.tril(np.ones((rows, cols)), cols - rows) > 0
        S = np.where(mask, S, -np.inf)
        e = np.exp(S - S.max(-1, keepdims=True))
        out[...] = e / e.sum(-1, keepdims=True)

    def mas_softmax_causal_backward(self, p, dp, ds, mats, rows, cols, scale):
        P = _f32(_addr(p), mats * rows * cols).reshape(mats, rows, cols).astype(np.float64)
        dP = _f32(_addr(dp), mats * rows * cols).reshape(mats, rows, cols).astype(np.float64)
        out = _f32(_addr(ds), mats * rows * cols).reshape(mats, rows, cols)
        mask = np.tril(np.ones((rows, cols)), cols - rows) > 0
        dPv = np.where(mask, dP, 0.0)                                   # never reads beyond the visible columns
        out[...] = np.where(mask, P * (dPv - (dPv * P).sum(-1, keepdims=True)) * scale, 0.0)

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


def _emu_more():
    """The remaining entries of the token transformer (training and cached decoding), same conventions."""

    def vec(p, n):
        return _f32(_addr(p), n)

    def ln(x, g, b, eps):
        m = x.mean(-1, keepdims=True)
        v = ((x - m) ** 2).mean(-1, keepdims=True)
        return (x - m) / np.sqrt(v + eps) * g + b, m[:, 0], 1.0 / np.sqrt(v[:, 0] + eps)

    def gelu(v):
        return 0.5 * v * (1.0 + np.tanh(0.7978845608028654 * v * (1.0 + 0.044715 * v * v)))

    def mas_gemm(self, A, B, C, M, N, K, batch, lda, ldb, ldc, sa, sb, sc, ta, tb, alpha, bias, residual, impl):
        for i in range(batch):
            Am = _mat(_addr(A) + 4 * i * sa, M, K, lda, bool(ta)).astype(np.float64)
            Bm = _mat(_addr(B) + 4 * i * sb, N, K, ldb, not bool(tb)).astype(np.float64)
            out = alpha * (Am @ Bm.T)
            if bias is not None:
                out = out + vec(bias, N).astype(np.float64)[None, :]
            if residual is not None:
                out = out + _mat(_addr(residual) + 4 * i * sc, M, N, ldc, False)
            _mat(_addr(C) + 4 * i * sc, M, N, ldc, False)[...] = out

    def mas_layernorm_forward(self, x, gamma, beta, residual, y, mean, rstd, R, H, eps):
        X = vec(x, R * H).reshape(R, H).astype(np.float64)
        o, m, rs = ln(X, vec(gamma, H).astype(np.float64), vec(beta, H).astype(np.float64), eps)
        if residual is not None:
            o = o + vec(residual, R * H).reshape(R, H)
        vec(y, R * H).reshape(R, H)[...] = o
        vec(mean, R)[...] = m
        vec(rstd, R)[...] = rs

    def mas_layernorm_backward(self, dy, x, mean, rstd, gamma, dx, dgamma, dbeta, R, H, ws, ws_bytes):
        D = vec(dy, R * H).reshape(R, H).astype(np.float64)
        X = vec(x, R * H).reshape(R, H).astype(np.float64)
        xh = (X - vec(mean, R).astype(np.float64)[:, None]) * vec(rstd, R).astype(np.float64)[:, None]
        g = D * vec(gamma, H).astype(np.float64)
        vec(dx, R * H).reshape(R, H)[...] = vec(rstd, R).astype(np.float64)[:, None] * (
            g - g.mean(-1, keepdims=True) - xh * (g * xh).mean(-1, keepdims=True))
        if dgamma is not None:
            vec(dgamma, H)[...] = (D * xh).sum(0)
            vec(dbeta, H)[...] = D.sum(0)

    def mas_gelu_forward(self, x, y, n):
        vec(y, n)[...] = gelu(vec(x, n).astype(np.float64))

    def mas_gelu_backward(self, dy, x, dx, n):
        v = vec(x, n).astype(np.float64)
        t = np.tanh(0.7978845608028654 * v * (1.0 + 0.044715 * v * v))
        du = 0.7978845608028654 * (1.0 + 3 * 0.044715 * v * v)
        vec(dx, n)[...] = vec(dy, n) * (0.5 * (1 + t) + 0.5 * v * (1 - t * t) * du)

    def mas_embed3_forward(self, t0, id0, t1, id1, t2, id2, out, R, H, seg, total, off):
        i0 = _i64(_addr(id0), R)
        i1 = _i64(_addr(id1), seg) if t1 is not None else None
        i2 = _i64(_addr(id2), seg) if t2 is not None else None
        rows = max(R // seg, 1) * total
        O = vec(out, rows * H).reshape(rows, H)
        for r in range(R):
            v = _f32(_addr(t0) + 4 * int(i0[r]) * H, H).astype(np.float64)
            if i1 is not None:
                v = v + _f32(_addr(t1) + 4 * int(i1[r % seg]) * H, H)
            if i2 is not None:
                v = v + _f32(_addr(t2) + 4 * int(i2[r % seg]) * H, H)
            O[(r // seg) * total + off + r % seg] = v

    def mas_embed3_backward(self, dout, id0, d0, id1, d1, id2, d2, R, H, seg, total, off):
        i0 = _i64(_addr(id0), R)
        rows = max(R // seg, 1) * total
        D = vec(dout, rows * H).reshape(rows, H)
        for r in range(R):
            g = D[(r // seg) * total + off + r % seg]
            _f32(_addr(d0) + 4 * int(i0[r]) * H, H)[...] += g
            if d1 is not None:
                _f32(_addr(d1) + 4 * int(_i64(_addr(id1), seg)[r % seg]) * H, H)[...] += g
            if d2 is not None:
                _f32(_addr(d2) + 4 * int(_i64(_addr(id2), seg)[r % seg]) * H, H)[...] += g

    def mas_conv1x1_wgrad(self, x, ldx, dy, ldy, M, cin, cout, dw, db, impl, ws, ws_bytes):
        X = _mat(_addr(x), M, cin, ldx, False).astype(np.float64)
        D = _mat(_addr(dy), M, cout, ldy, False).astype(np.float64)
        vec(dw, cout * cin).reshape(cout, cin)[...] = D.T @ X
        if db is not None:
            vec(db, cout)[...] = D.sum(0)

    def mas_linear_small(self, x, ldx, W, bias, y, ldy, R, N, K, act):
        X = _mat(_addr(x), R, K, ldx, False).astype(np.float64)
        Wm = vec(W, N * K).reshape(N, K).astype(np.float64)
        o = X @ Wm.T
        if bias is not None:
            o = o + vec(bias, N)[None, :]
        _mat(_addr(y), R, N, ldy, False)[...] = gelu(o) if act == 1 else o

    def mas_kv_append(self, qkv, R, T, heads, hd, kc, vc, Tmax, pos0):
        H = heads * hd
        Q = vec(qkv, R * T * 3 * H).reshape(R, T, 3, heads, hd)
        K_ = vec(kc, R * heads * Tmax * hd).reshape(R, heads, Tmax, hd)
        V_ = vec(vc, R * heads * Tmax * hd).reshape(R, heads, Tmax, hd)
        K_[:, :, pos0:pos0 + T] = Q[:, :, 1].transpose(0, 2, 1, 3)
        V_[:, :, pos0:pos0 + T] = Q[:, :, 2].transpose(0, 2, 1, 3)

    def mas_attn_decode(self, qkv, kc, vc, ctx, R, heads, hd, Tmax, length):
        H = heads * hd
        q = vec(qkv, R * 3 * H).reshape(R, 3, heads, hd)[:, 0].astype(np.float64)
        K_ = vec(kc, R * heads * Tmax * hd).reshape(R, heads, Tmax, hd)[:, :, :length].astype(np.float64)
        V_ = vec(vc, R * heads * Tmax * hd).reshape(R, heads, Tmax, hd)[:, :, :length].astype(np.float64)
        s = np.einsum("rhd,rhtd->rht", q, K_) / np.sqrt(hd)
        p = np.exp(s - s.max(-1, keepdims=True))
        p = p / p.sum(-1, keepdims=True)
        vec(ctx, R * H).reshape(R, heads, hd)[...] = np.einsum("rht,rhtd->rhd", p, V_)

    def mas_attn_decode_append(self, qkv, kc, vc, ctx, R, heads, hd, Tmax, pos):
        self.mas_kv_append(qkv, R, 1, heads, hd, kc, vc, Tmax, pos)
        self.mas_attn_decode(qkv, kc, vc, ctx, R, heads, hd, Tmax, pos + 1)

    def mas_layernorm2_forward(self, x, g1, b1, residual, y1, g2, b2, y2, R, H, eps1, eps2):
        X = vec(x, R * H).reshape(R, H).astype(np.float64)
        o, _, _ = ln(X, vec(g1, H).astype(np.float64), vec(b1, H).astype(np.float64), eps1)
        if residual is not None:
            o = o + vec(residual, R * H).reshape(R, H)
        vec(y1, R * H).reshape(R, H)[...] = o
        o2, _, _ = ln(vec(y1, R * H).reshape(R, H).astype(np.float64), vec(g2, H).astype(np.float64), vec(b2, H).astype(np.float64), eps2)
        vec(y2, R * H).reshape(R, H)[...] = o2

    def mas_cfg_mix(self, cond, uncond, out, n, scale):
        u = vec(uncond, n).astype(np.float64)
        vec(out, n)[...] = u + scale * (vec(cond, n) - u)

    def mas_sample_topk(self, logits, ld, R, V, temperature, top_k, u, tokens):
        Z = _mat(_addr(logits), R, V, ld, False).astype(np.float64) / temperature
        uu = vec(u, R)
        out = _i64(_addr(tokens), R)
        for r in range(R):
            z = Z[r].copy()
            if 0 < top_k < V:
                z[z < np.sort(z)[-top_k]] = -np.inf
            p = np.exp(z - z.max())
            c = np.cumsum(p)
            out[r] = min(int(np.searchsorted(c, uu[r] * c[-1], side="right")), V - 1)

    # ---- fp16 Linear path (csrc/gemm_tma.cu): real fp16 rounding of the operand copies, power-of-two scales from max|.|
    packs = {}

    def f16(p, n):
        return np.ctypeslib.as_array((ctypes.c_uint16 * n).from_address(_addr(p))).view(np.float16)

    def scale_of(amax):
        if amax is None:
            return 1.0
        a = float(vec(amax, 1)[0])
        return 1.0 if not np.isfinite(a) or a <= 0 else 2.0 ** (14 - int(np.floor(np.log2(a))))

    def mas_amax(self, x, n, out):
        vec(out, 1)[0] = np.abs(vec(x, n)).max()

    def mas_to_half(self, x, y, n, amax):
        f16(y, n)[...] = (vec(x, n).astype(np.float64) * scale_of(amax)).astype(np.float16)

    def mas_pack_gemm_tc16(self, w, out, N, K, transpose):
        W = vec(w, N * K).reshape(N, K).astype(np.float16).astype(np.float64)          # operand rounding of the weights
        packs[_addr(out)] = W.T.copy() if transpose else W                             # rows = output features of the product

    def mas_gemm_rows_f16(self, x16, M, K, wpk, y, ldy, N, bias, residual, x_amax, alpha):
        X = f16(x16, M * K).reshape(M, K).astype(np.float64) / scale_of(x_amax)
        Wp = packs[_addr(wpk)]
        assert Wp.shape == (N, K), (Wp.shape, N, K)
        o = alpha * (X @ Wp.T)
        if bias is not None:
            o = o + vec(bias, N)[None, :]
        if residual is not None:
            o = o + _mat(_addr(residual), M, N, ldy, False)
        _mat(_addr(y), M, N, ldy, False)[...] = o

    def mas_wgrad_rows_f16(self, x16, dy16, M, N, K, dw, db, x_amax, dy_amax, ws, ws_bytes):
        X = f16(x16, M * K).reshape(M, K).astype(np.float64) / scale_of(x_amax)
        D = f16(dy16, M * N).reshape(M, N).astype(np.float64) / scale_of(dy_amax)
        vec(dw, N * K).reshape(N, K)[...] = D.T @ X
        if db is not None:
            vec(db, N)[...] = D.sum(0)

    for k, v in list(locals().items()):
        if k.startswith("mas_"):
            setattr(Emu, k, v)


_emu_more()


@pytest.fixture
def emu(monkeypatch):
    from mas_b200 import ops
    e = Emu()
    monkeypatch.setattr(ops.L, "call", e)
    monkeypatch.setattr(ops, "_need_cuda", lambda x: None)
    monkeypatch.setattr(ops, "_tc_on", lambda: False)
    monkeypatch.setattr(ops.L, "query", lambda name, *a: 256)
    monkeypatch.setattr(torch.cuda, "graph_pool_handle", lambda: None)
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


def _tiny_model():
    import os
    from conftest import GOLDEN
    from models.transformer import MakeAScene
    g = torch.load(os.path.join(GOLDEN, "transformer_tiny.pt"), weights_only=False)
    m = MakeAScene(**g["cfg"])
    m.load_state_dict(g["state_dict"])
    m.device = torch.device("cpu")
    return m, g


def test_make_a_scene_host_logic_against_reference_fixture(emu):
    """The whole drop-in module above an emulated C-ABI reproduces the REAL reference's logits, loss and gradients
    (tests/golden/transformer_tiny.pt): everything models/transformer.py and the autograd units do between kernel calls."""
    from conftest import rel_err
    m, g = _tiny_model()
    logits = m(g["text"], g["seg"], g["img"])
    assert rel_err(logits, g["logits"]) < 1e-5
    loss = m.loss(g["text"], g["seg"], g["img"])
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-5
    loss.backward()
    named = dict(m.named_parameters())
    for k, gv in g["grads"].items():
        assert rel_err(named[k].grad, gv) < 1e-4, k
    assert "mas_ce_forward" in emu.names and "mas_ce_backward" in emu.names


def test_generate_host_logic_against_reference_fixture(emu):
    """KV-cached decoding (prefill, fused append + attention, chained LayerNorm pairs, guidance mix, sampler entry) above the
    emulated C-ABI: teacher-forced logits == the reference's non-cached logits; launches per decoded token as designed."""
    from conftest import rel_err
    m, g = _tiny_model()
    m.eval()
    toks, lg = m.generate(g["text"], g["seg"], img_tokens=g["img"], return_logits=True)
    assert torch.equal(toks, g["img"]) and rel_err(lg, g["logits"]) < 1e-5
    layers = len(m.transformer.layers)
    steps = m.image_length - 1
    assert emu.names.count("mas_attn_decode_append") == layers * steps
    assert emu.names.count("mas_layernorm2_forward") == 2 * layers * steps
    assert emu.names.count("mas_kv_append") == layers                       # the prefill only
    # guidance: logits = uncond + s (cond - uncond), both streams through one decode pass
    with torch.no_grad():
        cond = m(g["text"], g["seg"], g["img"])
        uncond = m(torch.zeros_like(g["text"]), g["seg"], g["img"])
    _, lg2 = m.generate(g["text"], g["seg"], guidance_scale=2.5, img_tokens=g["img"], return_logits=True)
    assert rel_err(lg2, uncond + 2.5 * (cond - uncond)) < 1e-5
    # sampling: seeded, top-k respected
    gen = torch.Generator().manual_seed(5)
    a, la = m.generate(g["text"], g["seg"], temperature=0.9, top_k=3, generator=gen, return_logits=True)
    kth = la.topk(3, dim=-1).values[..., -1]
    assert bool((la.gather(-1, a.unsqueeze(-1)).squeeze(-1) >= kth).all())
    assert "mas_sample_topk" in emu.names


def test_linear_fp16_path_host_logic_against_reference_fixture(emu, monkeypatch):
    """LinearFn's fp16 route (one conversion of the input kept for the weight gradient, one of the output gradient for both
    gradients, weight images cached per parameter version) above the emulated entries - with real fp16 operand rounding - stays
    within the production tolerance of the REAL reference's logits and gradients (tests/golden/transformer_wide.pt)."""
    import os
    from conftest import GOLDEN, rel_err
    from mas_b200 import ops
    from models.transformer import MakeAScene
    monkeypatch.setattr(ops, "f16_operands", lambda: True)
    monkeypatch.setenv("MAS_LINEAR_F16", "1")
    g = torch.load(os.path.join(GOLDEN, "transformer_wide.pt"), weights_only=False)
    m = MakeAScene(**g["cfg"])
    m.load_state_dict(g["state_dict"])
    m.device = torch.device("cpu")
    loss = m.loss(g["text"], g["seg"], g["img"])
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-3 * max(1.0, float(g["loss"]))
    loss.backward()
    assert emu.names.count("mas_gemm_rows_f16") == 2 * 4 * g["cfg"]["num_layers"] + 2       # forward + data gradient of every Linear
    assert emu.names.count("mas_wgrad_rows_f16") == 4 * g["cfg"]["num_layers"] + 1
    n_lin = 4 * g["cfg"]["num_layers"] + 1
    assert emu.names.count("mas_to_half") == 2 * n_lin and emu.names.count("mas_pack_gemm_tc16") == 2 * n_lin
    named = dict(m.named_parameters())
    for k, gv in g["grads"].items():
        assert rel_err(named[k].grad, gv) < 5e-3, k
    # a second step without touching the weights packs nothing; an in-place update repacks
    before = emu.names.count("mas_pack_gemm_tc16")
    m.zero_grad()
    m.loss(g["text"], g["seg"], g["img"]).backward()
    assert emu.names.count("mas_pack_gemm_tc16") == before
    with torch.no_grad():
        m.to_logits[1].weight.add_(0.0)
    m.loss(g["text"], g["seg"], g["img"]).backward()
    assert emu.names.count("mas_pack_gemm_tc16") == before + 2
