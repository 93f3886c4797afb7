"""Convolution kernels in isolation, timed with CUDA events (and a hook for ncu: `micro_conv.py one <variant>`).
Variants: {tf32,f16} x {plain, pro (GroupNorm+SiLU prologue + statistics epilogue)} forward, dgrad, wgrad."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "make-a-scene_b200")]
import torch  # noqa: E402
from mas_b200 import _lib as L, ops  # noqa: E402

dev = torch.device("cuda:0")


def bench(fn, it=6, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def variants(B, C, H, Co=None):
    Co = Co or C
    x = torch.randn(B, C, H, H, device=dev).contiguous(memory_format=torch.channels_last)
    dy = (torch.randn(B, Co, H, H, device=dev) * 1e-6).contiguous(memory_format=torch.channels_last)
    w = torch.randn(Co, C, 3, 3, device=dev) * 0.03
    b = torch.zeros(Co, device=dev)
    g, be = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    m, r = ops.gn_stats(x)
    tab = ops.gn_table(m, r, g, be, B, C)
    xa, dya = ops.amax(x), ops.amax(dy)
    out = {}
    for fmt in ("tf32", "f16"):
        def f_plain(fmt=fmt):
            ops.set_operand_format(fmt)
            ops.conv3x3_raw(x, w, b, None, L.CONV_S1, x_amax=xa)

        def f_pro(fmt=fmt):
            ops.set_operand_format(fmt)
            ops.conv3x3_raw(x, w, b, x, L.CONV_S1, table=tab, want_stats=True) if C == Co else ops.conv3x3_raw(x, w, b, None, L.CONV_S1, table=tab, want_stats=True)

        def f_dgrad(fmt=fmt):
            ops.set_operand_format(fmt)
            ops.conv3x3_dgrad_raw(dy, w, L.CONV_S1, dya)

        def f_wgrad(fmt=fmt):
            ops.set_operand_format(fmt)
            ops.conv3x3_wgrad_raw(x, dy, Co, C, L.CONV_S1, dy_amax=dya)
        out[fmt + "_plain"], out[fmt + "_pro"], out[fmt + "_dgrad"], out[fmt + "_wgrad"] = f_plain, f_pro, f_dgrad, f_wgrad
    out["amax"] = lambda: ops.amax(x)
    if C % 64 == 0 and Co % 128 == 0:
        ops.set_operand_format("f16")
        x16 = ops.to_half(x)
        dy16 = ops.to_half(dy, dya)
        out["tma_plain"] = lambda: ops.conv3x3_h_raw(x16, w, b, None)
        out["tma_stats"] = lambda: ops.conv3x3_h_raw(x16, w, b, x if C == Co else None, want_stats=True)
        out["tma_dgrad"] = lambda: ops.conv3x3_h_raw(dy16, w, None, None, transpose=True, x_amax=dya)
        out["wgrad16"] = lambda: ops.conv3x3_wgrad_raw(x16, dy16, Co, C, L.CONV_S1, dy_amax=dya)   # both operands as fp16 shadows
        out["gn_apply16"] = lambda: ops.gn_apply_f16(x, m, r, g, be, True)
        out["to_half"] = lambda: ops.to_half(x)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "one":
        v = variants(32, 128, 256)
        for name in sys.argv[2:]:
            v[name](); v[name]()
        torch.cuda.synchronize()
        print("done", sys.argv[2:])
        sys.exit(0)
    for (B, C, H, Co) in ((32, 128, 256, 128), (32, 256, 64, 256), (32, 512, 32, 512), (32, 512, 16, 512)):
        v = variants(B, C, H, Co)
        gf = 2.0 * B * H * H * C * Co * 9 / 1e9
        print("shape B%d C%d->%d @%d  (%.1f GFLOP)" % (B, C, Co, H, gf))
        for k, fn in v.items():
            ms = bench(fn)
            print("  %-12s %8.3f ms  %7.1f TFLOP/s" % (k, ms, gf / ms if k not in ("amax", "gn_apply16", "to_half") else 0))
        del v
        torch.cuda.empty_cache()
