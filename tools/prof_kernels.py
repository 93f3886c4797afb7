"""Runs the dominant kernels in isolation (for ncu captures): conv3x3 128->128 @256^2 x32 fprop / wgrad on the tcgen05
path, GroupNorm kernels, VQ.  Usage: python tools/prof_kernels.py [conv|wgrad|gn|vq|attn|all]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "make-a-scene_b200")]
import torch  # noqa: E402
from mas_b200 import _lib as L, ops  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(os.environ.get("PROF_BATCH", "32"))
dev = torch.device("cuda:0")
big = what in ("conv", "wgrad", "gn", "all")
x = torch.randn(B if big else 1, 128, 256, 256, device=dev).contiguous(memory_format=torch.channels_last)
dy = torch.randn(B if big else 1, 128, 256, 256, device=dev).contiguous(memory_format=torch.channels_last)
w = torch.randn(128, 128, 3, 3, device=dev) * 0.03
b = torch.zeros(128, device=dev)
g, be = torch.ones(128, device=dev), torch.zeros(128, device=dev)
for _ in range(2):
    if what in ("conv", "all"):
        y = ops.conv3x3_raw(x, w, b, None, L.CONV_S1)
    if what in ("wgrad", "all"):
        dw, db = ops.conv3x3_wgrad_raw(x, dy, 128, 128, L.CONV_S1)
    if what in ("gn", "all"):
        m, r = ops.gn_stats(x)
        a = ops.gn_apply(x, m, r, g, be, True)
        dx, dg, dbb = ops.gn_backward(dy, x, m, r, g, be, True)
    if what in ("vq", "all"):
        z = torch.randn(32, 256, 16, 16, device=dev).contiguous(memory_format=torch.channels_last)
        E = torch.randn(8192, 256, device=dev)
        ops.vq_select_path(True)
        ops.VQFn.apply(z, E, 0.25)
        ops.vq_select_path(False)
        ops.VQFn.apply(z, E, 0.25)
        ops.vq_select_path(True)
    if what in ("attn",):
        from models import modules as M
        blk = M.AttnBlock(512).to(dev)
        xa = torch.randn(B, 512, 16, 16, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        blk(xa).backward(torch.randn(B, 512, 16, 16, device=dev).contiguous(memory_format=torch.channels_last))
torch.cuda.synchronize()
print("done", what)
