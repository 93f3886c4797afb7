"""Runs conv_in (3 -> 128) and conv_out (128 -> 3) forward + backward at 256^2 in isolation (for ncu captures of the
edge kernels).  Usage: python tools/prof_edge.py   (PROF_BATCH overrides the batch of 32)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "make-a-scene_b200")]
import torch  # noqa: E402
from mas_b200 import _lib as L, ops  # noqa: E402

B = int(os.environ.get("PROF_BATCH", "32"))
dev = torch.device("cuda:0")
img = torch.rand(B, 3, 256, 256, device=dev, requires_grad=True)
w_in = (torch.randn(128, 3, 3, 3, device=dev) * 0.1).requires_grad_(True)
b_in = torch.zeros(128, device=dev, requires_grad=True)
a = torch.randn(B, 128, 256, 256, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
w_out = (torch.randn(3, 128, 3, 3, device=dev) * 0.03).requires_grad_(True)
b_out = torch.zeros(3, device=dev, requires_grad=True)
for _ in range(2):
    y = ops.Conv3x3Fn.apply(img, w_in, b_in, None, L.CONV_S1, False)
    y.backward(torch.ones_like(y))
    r = ops.Conv3x3Fn.apply(a, w_out, b_out, None, L.CONV_S1, False)
    r.backward(torch.ones_like(r))
torch.cuda.synchronize()
print("done edge")
