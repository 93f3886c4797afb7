"""Address-reveal probe for kind::f16 K-major operands under the 128-byte swizzle: which half of shared memory does the tensor
core read for element (n, k) when the descriptor start address is moved by whole rows (128 B) and by K steps (32 B), and when the
8-row group stride (SBO) is not a multiple of the 1024-byte swizzle atom?  The TMA-fed convolution kernel relies on the answer:
the 3x3 taps are descriptor shifts over ONE staged halo whose image rows are 10 pixels (1280 B) apart.

Hypothesis checked here ("absolute"): byte address a = start + (n%8)*128 + (n/8)*SBO + k*2, then bits [4,7) ^= bits [7,10) of a.
"""
import sys
import os
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "make-a-scene_b200"))
from mas_b200 import _lib as L  # noqa: E402


def desc(lbo, sbo, layout_type, base_off=0):
    return ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (base_off << 49) | (layout_type << 61)


def idesc16(n):
    return (1 << 4) | ((n >> 3) << 17) | ((128 >> 4) << 24)


def predict(n, k, off, sbo):
    a = off + (n % 8) * 128 + (n // 8) * sbo + k * 2
    a ^= ((a >> 7) & 7) << 4
    return a // 2


CASES = [  # name, sbo, off, n_cols, base_off
    ("canonical", 1024, 0, 32, 0),
    ("k+16 (32 B)", 1024, 32, 16, 0),
    ("k+48 (96 B)", 1024, 96, 16, 0),
    ("row+1 (128 B)", 1024, 128, 16, 0),
    ("row+3, k+32", 1024, 3 * 128 + 64, 16, 0),
    ("sbo 1280", 1280, 0, 16, 0),
    ("sbo 1280 row+1", 1280, 128, 16, 0),
    ("sbo 1280 row+11 k+16", 1280, 11 * 128 + 32, 16, 0),
    ("row+1 base_off 1", 1024, 128, 16, 1),
    ("sbo 1280 row+1 base_off 1", 1280, 128, 16, 1),
]

dev = torch.device("cuda:0")
ok_all = True
for name, sbo, off, ncols, base in CASES:
    D = torch.full((128, 32), float("nan"), device=dev)
    L.call("mas_tc_probe16", D, desc(16, sbo, 2, base), idesc16(ncols), off)
    torch.cuda.synchronize()
    o = D[:16].t().cpu()
    bad = 0
    rows = []
    for n in range(ncols):
        got = [int(v) if v == v else -1 for v in o[n].tolist()]
        want = [predict(n, k, off, sbo) for k in range(16)]
        want = [w if w < 2048 else None for w in want]
        miss = sum(1 for g, w in zip(got, want) if w is not None and g != w)
        bad += miss
        rows.append((n, got, want, miss))
    print(f"SW128 {name}: sbo={sbo} off={off} base_off={base} -> {'MATCH' if bad == 0 else 'MISMATCH (%d)' % bad}")
    if bad:
        ok_all = ok_all and base != 0
        for n, got, want, miss in rows[:12]:
            print("   n=%2d got %s" % (n, got))
            print("        exp %s" % (want,))
print("ABSOLUTE-ADDRESS HYPOTHESIS", "HOLDS" if ok_all else "FAILS")
