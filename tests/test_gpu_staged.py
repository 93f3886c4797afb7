"""Staged kernels (built and exported, selected by NO module path until validated on a B200) - skipped unless
MAS_EXPERIMENTAL=1:

  * csrc/contract_tc3.cu: fp32-accurate 3xTF32 batched GEMM on tcgen05, reachable only through mas_gemm(impl=MAS_IMPL_TC3)
        MAS_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_staged.py -m gpu -k tc3
  * csrc/contract_tc2.cu: the 3x3 convolution forward / data-gradient kernel with cta_group::2 CTA pairs, selected
    process-wide by MAS_CONV_2CTA=1 (read once, at the first convolution)
        MAS_EXPERIMENTAL=1 MAS_CONV_2CTA=1 timeout 300 python -m pytest tests -m gpu      # the whole parity suite on it
    (run under `timeout`: a wrong barrier protocol hangs the kernel)."""
import os

import pytest
import torch

from conftest import rel_err

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MAS_EXPERIMENTAL") != "1", reason="staged kernel: set MAS_EXPERIMENTAL=1")]


@pytest.mark.parametrize("M,N,K,batch,ta,tb", [(256, 256, 512, 4, 0, 1),     # S = Q K^T
                                               (256, 512, 256, 3, 0, 0),     # O = P V
                                               (256, 512, 256, 2, 1, 0),     # dV = P^T dO
                                               (256, 256, 512, 2, 0, 1),     # dP = dO V^T
                                               (640, 640, 64, 5, 0, 1),      # transformer head: S
                                               (640, 64, 640, 5, 0, 0),      # transformer head: P V (N tile 64)
                                               (640, 64, 640, 2, 1, 0),      # transformer head: dV = P^T dO
                                               (100, 128, 32, 1, 0, 1), (100, 128, 64, 2, 1, 0), (72, 192, 96, 2, 0, 0)])   # ragged M
def test_gemm_tc3_vs_fp64(M, N, K, batch, ta, tb):
    from mas_b200 import _lib as L
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + N + K + batch)
    A = torch.randn((batch, K, M) if ta else (batch, M, K), generator=g)
    B = torch.randn((batch, N, K) if tb else (batch, K, N), generator=g)
    opA = A.transpose(1, 2) if ta else A
    opB = B if tb else B.transpose(1, 2)
    ref = 0.37 * (opA.double() @ opB.double().transpose(1, 2))
    Ad, Bd = A.to(dev), B.to(dev)
    C = torch.empty(batch, M, N, device=dev)
    lda = M if ta else K
    ldb = K if tb else N
    L.call("mas_gemm", Ad, Bd, C, M, N, K, batch, lda, ldb, N, A[0].numel(), B[0].numel(), M * N, ta, tb, 0.37, None, None, L.IMPL_TC3)
    # fp32-level accuracy (a single TF32 pass would sit at ~5e-4)
    assert rel_err(C, ref.float()) < 5e-6


@pytest.mark.skipif(os.environ.get("MAS_CONV_2CTA") != "1", reason="run with MAS_CONV_2CTA=1 (selected once per process)")
@pytest.mark.parametrize("cin,cout,h,w,n,mode", [(128, 128, 32, 32, 2, "s1"), (128, 128, 256, 256, 2, "s1"), (256, 128, 64, 64, 3, "s1"),
                                                (512, 512, 16, 16, 32, "s1"), (128, 128, 16, 16, 2, "up"), (64, 128, 16, 8, 1, "s1")])
def test_conv3x3_cta_pair_vs_fp32(cin, cout, h, w, n, mode):
    """cta_group::2 convolution (odd tile counts leave the last pair half empty) against the exact-fp32 SIMT kernel."""
    from mas_b200 import _lib as L, ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(n, cin, h, w, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).to(dev)
    b = torch.randn(cout, generator=g).to(dev)
    md = L.CONV_S1 if mode == "s1" else L.CONV_UP
    ops.set_impl(L.IMPL_AUTO)
    y = ops.conv3x3_raw(x, wt, b, None, md)
    ops.set_impl(L.IMPL_SIMT)
    try:
        ref = ops.conv3x3_raw(x, wt, b, None, md)
    finally:
        ops.set_impl(L.IMPL_AUTO)
    assert rel_err(y, ref) < 1e-3
