"""Address-reveal probe: kind::f16 MN-major operand under the 128-byte swizzle (rows = K index, 64 MN elements per 128-byte
row), with a K-group (8 rows) stride that is not a multiple of the swizzle atom and row-shifted starts - the layout a
channels-last activation halo [pixel][64 channels] has when the reduction runs over pixels (weight gradient).

Hypothesis: byte address a = start + (k%8)*128 + (k/8)*KG + n*2 with KG taken from the SBO field (variant A) or the LBO field
(variant B); bits [4,7) ^= bits [7,10) of a."""
import os
import sys
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "make-a-scene_b200"))
from mas_b200 import _lib as L  # noqa: E402


def desc(lbo, sbo, layout_type):
    return ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (layout_type << 61)


def idesc16(n, b_mn):
    return (1 << 4) | ((1 << 16) if b_mn else 0) | ((n >> 3) << 17) | ((128 >> 4) << 24)


def predict(n, k, off, kg):
    a = off + (k % 8) * 128 + (k // 8) * kg + n * 2
    a ^= ((a >> 7) & 7) << 4
    return a // 2


dev = torch.device("cuda:0")
CASES = [  # name, lbo, sbo, off, kg expected
    ("A: sbo=1024 (canonical)", 16, 1024, 0, 1024),
    ("A: sbo=1280", 16, 1280, 0, 1280),
    ("A: sbo=1280 off=128", 16, 1280, 128, 1280),
    ("A: sbo=1280 off=1280+256", 16, 1280, 1280 + 256, 1280),
    ("B: lbo=1024 sbo=16", 1024, 16, 0, 1024),
    ("B: lbo=1280 sbo=16", 1280, 16, 0, 1280),
    ("B: lbo=1280 sbo=4096 off=128", 1280, 4096, 128, 1280),
]
for name, lbo, sbo, off, kg in CASES:
    D = torch.full((128, 32), float("nan"), device=dev)
    L.call("mas_tc_probe16", D, desc(lbo, sbo, 2), idesc16(32, 1), off)
    torch.cuda.synchronize()
    o = D[:16].t().cpu()      # [n][k]
    bad = 0
    for n in range(32):
        for k in range(16):
            w = predict(n, k, off, kg)
            g = o[n][k].item()
            if w < 2048 and (g != g or int(g) != w):
                bad += 1
    print("MN-SW128 %s: lbo=%d sbo=%d off=%d -> %s" % (name, lbo, sbo, off, "MATCH" if bad == 0 else "MISMATCH (%d)" % bad))
    if bad:
        for n in (0, 1, 8, 9, 31):
            print("   n=%2d got %s" % (n, [int(v) if v == v else -1 for v in o[n].tolist()]))
            print("        exp %s" % [predict(n, k, off, kg) for k in range(16)])
