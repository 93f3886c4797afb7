"""TMA-fed fp16 convolution kernel (csrc/conv_tma.cu) against the register-staged fp16 kernel (same operand rounding, same
accumulation order) and against F.conv2d in fp32; the fp16 shadow writers it is fed by."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


SHAPES = [  # n, cin, cout, h, w
    (2, 128, 128, 32, 32),
    (1, 64, 256, 16, 16),      # two output-channel tiles, one 64-channel chunk
    (3, 256, 128, 16, 8),      # image narrower than the 10-pixel halo box, odd tile count
    (1, 128, 128, 16, 24),     # three tiles: the last pair is half empty
    (2, 512, 512, 16, 16),
    (1, 128, 160, 32, 16),     # padded output channels (stored through Cstore)
    (1, 128, 128, 64, 8),      # 32 x 8 tiles (N = 256 MMAs), two per image column
]


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_tma_matches_staged_kernel_and_fp32(shape):
    from mas_b200 import _lib as L, ops
    n, cin, cout, h, w = shape
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(7)
    x = _cl(torch.randn(n, cin, h, w, generator=g).to(dev))
    wt = (torch.randn(cout, cin, 3, 3, generator=g) * (1.0 / (3.0 * cin ** 0.5))).to(dev)
    b = torch.randn(cout, generator=g).to(dev)
    ops.set_operand_format("f16")
    x16 = ops.to_half(x)
    assert torch.equal(x16, x.half())
    if cout % 128 == 0:
        y = ops.conv3x3_h_raw(x16, wt, b, None)
    else:
        # padded head: weights / bias padded to the 128-wide tile by the caller, output stores the real channels
        cp = (cout + 127) // 128 * 128
        wp = torch.zeros(cp, cin, 3, 3, device=dev); wp[:cout] = wt
        bp = torch.zeros(cp, device=dev); bp[:cout] = b
        y = ops.empty_nhwc(n, cout, h, w, x)
        pk = torch.empty(9 * cp * cin, dtype=torch.float16, device=dev)
        L.call("mas_pack_conv3x3_tc16", wp, pk, None, cp, cin, 0)
        L.call("mas_conv3x3_fprop_tc16h", x16, L.t4(x16), pk, bp, None, y, L.t4(y), None, None)
    ref = F.conv2d(x16.float(), wt.half().float(), b, padding=1)
    err = (y - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-4, err          # same fp16-rounded operands, fp32 accumulation: only the summation differs
    ref32 = F.conv2d(x, wt, b, padding=1)
    assert (y - ref32).abs().max().item() / ref32.abs().max().item() < 3e-3


def test_conv_tma_scaled_shadow_residual_and_statistics():
    """Gradient-like magnitudes (1e-7) through a scaled shadow; residual and GroupNorm-statistics epilogues."""
    from mas_b200 import _lib as L, ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(11)
    n, c, h, w = 2, 128, 32, 32
    x = _cl((torch.randn(n, c, h, w, generator=g) * 3e-7).to(dev))
    wt = (torch.randn(c, c, 3, 3, generator=g) * 0.03).to(dev)
    res = _cl(torch.randn(n, c, h, w, generator=g).to(dev) * 1e-6)
    ops.set_operand_format("f16")
    am = ops.amax(x)
    x16 = ops.to_half(x, am)
    assert x16.float().abs().max().item() >= 2.0 ** 14
    y, st = ops.conv3x3_h_raw(x16, wt, None, res, want_stats=True, x_amax=am)
    y_old = ops.conv3x3_raw(x, wt, None, res, L.CONV_S1, x_amax=am)
    # same operand rounding and accumulation order as the register-staged kernel
    print("bit-identical to the staged kernel:", torch.equal(y, y_old))
    assert (y - y_old).abs().max().item() <= 1e-6 * y_old.abs().max().item()
    ref = F.conv2d(x, wt, None, padding=1) + res
    assert (y - ref).abs().max().item() / ref.abs().max().item() < 3e-3
    m, r = ops.gn_stats(y)
    assert torch.allclose(st[0], m, rtol=1e-4, atol=1e-9) and torch.allclose(st[1], r, rtol=1e-4)
    # data-gradient operand (transposed, flipped taps)
    d = ops.conv3x3_h_raw(x16, wt, None, None, transpose=True, x_amax=am)
    dref = F.conv_transpose2d(x, wt, padding=1)
    assert (d - dref).abs().max().item() / dref.abs().max().item() < 3e-3


def test_gn_apply_f16_shadow():
    from mas_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(3)
    x = _cl((torch.randn(2, 128, 16, 24, generator=g) * 2 + 0.5).to(dev))
    gamma, beta = torch.randn(128, generator=g).to(dev), torch.randn(128, generator=g).to(dev)
    m, r = ops.gn_stats(x)
    for silu in (True, False):
        a = ops.gn_apply(x, m, r, gamma, beta, silu)
        a16 = ops.gn_apply_f16(x, m, r, gamma, beta, silu)
        assert a16.dtype == torch.float16 and a16.stride() == a.stride()
        assert torch.equal(a16, a.half())


@pytest.mark.parametrize("with_add", [False, True])
def test_gn_backward_shadow_scale_is_a_rigorous_bound(with_add):
    """The fp16 shadow of dx is scaled from a bound on max|dx| that must (a) hold and (b) not waste the fp16 range."""
    import math
    from mas_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(5)
    ops.set_operand_format("f16")
    n, c, h, w = 2, 128, 32, 16
    x = _cl((torch.randn(n, c, h, w, generator=g) * 1.7 + 0.3).to(dev))
    x[0, 5, 3, 3] = 40.0                      # an outlier: large |xhat|
    dy = _cl((torch.randn(n, c, h, w, generator=g) * 2e-6).to(dev))
    add = _cl((torch.randn(n, c, h, w, generator=g) * 5e-6).to(dev)) if with_add else None
    gamma, beta = (torch.randn(c, generator=g) * 0.5 + 1).to(dev), (torch.randn(c, generator=g) * 0.1).to(dev)
    m, r = ops.gn_stats(x)
    dx, dg, db = ops.gn_backward(dy, x, m, r, gamma, beta, True, dx_add=add, shadow=True)
    dx_ref, dg_ref, db_ref = ops.gn_backward(dy, x, m, r, gamma, beta, True, dx_add=add)
    assert torch.equal(dx, dx_ref) and torch.equal(dg, dg_ref) and torch.equal(db, db_ref)
    sh = ops.shadow_of(dx)
    assert sh is not None
    dx16, bound = sh
    amax = dx.abs().max().item()
    b = bound.item()
    print("bound / amax = %.2f" % (b / amax))
    assert b >= amax and b / amax < 64.0
    s = 2.0 ** (14 - math.floor(math.log2(b)))
    want = (dx.double() * s).half()
    assert torch.equal(dx16, want.to(dx16.dtype))
    assert torch.isfinite(dx16.float()).all()
    assert ops.amax_of(dx).item() == pytest.approx(amax, rel=0, abs=0)


@pytest.mark.parametrize("shape", [(2, 64, 128, 32, 16), (2, 128, 128, 16, 24), (1, 256, 256, 16, 16), (3, 128, 160, 8, 8)])
def test_wgrad_from_fp16_shadows(shape):
    """Weight gradient from the two fp16 shadows (pure TMA + MMA kernel, MN-major activation operand under the 128-byte
    swizzle) against the register-staged kernel on the same operands and against autograd in fp32."""
    from mas_b200 import _lib as L, ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(13)
    ops.set_operand_format("f16")
    n, cin, cout, h, w = shape
    x = _cl(torch.randn(n, cin, h, w, generator=g).to(dev))
    dy = _cl((torch.randn(n, cout, h, w, generator=g) * 3e-6).to(dev))
    am = ops.amax(dy)
    bound = am * 1.7          # any magnitude >= max|dy| is a valid scale source
    dy16 = ops.to_half(dy, bound)
    x16 = ops.to_half(x)
    rows = (cout + 127) // 128 * 128           # the padded head: dw / dbias sized for the 128-wide tile
    dw, db = ops.conv3x3_wgrad_raw(x16, dy16, rows, cin, L.CONV_S1, dy_amax=bound)
    assert L.tc_launch_count() > 0
    xr = x.clone().requires_grad_(True)
    wr = torch.zeros(cout, cin, 3, 3, device=dev, requires_grad=True)
    br = torch.zeros(cout, device=dev, requires_grad=True)
    F.conv2d(xr, wr, br, padding=1).backward(dy)
    assert (dw[:cout] - wr.grad).abs().max().item() / wr.grad.abs().max().item() < 3e-3
    assert (db[:cout] - br.grad).abs().max().item() / br.grad.abs().max().item() < 1e-3
    if rows != cout:
        assert dw[cout:].abs().max().item() == 0.0 and db[cout:].abs().max().item() == 0.0
    else:
        # fp32 dy through the register-staged kernel with the same scale: same fp16 operands, different summation order
        dw0, db0 = ops.conv3x3_wgrad_raw(x16, dy, cout, cin, L.CONV_S1, dy_amax=bound)
        assert (dw - dw0).abs().max().item() <= 2e-5 * dw0.abs().max().item()
        assert (db0 - br.grad).abs().max().item() / br.grad.abs().max().item() < 1e-5
