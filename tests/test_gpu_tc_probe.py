"""GPU: pins the tcgen05 operand conventions (shared-memory descriptor majorness, A-from-TMEM) with a one-MMA probe."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a_src,b_layout", [(0, 0), (1, 0), (0, 1), (1, 1), (0, 2), (1, 2)])
def test_single_mma_conventions(a_src, b_layout):
    from mas_b200 import _lib as L
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    A = torch.randn(128, 8, generator=g).to(dev)
    B = torch.randn(32, 8, generator=g).to(dev)
    D = torch.full((128, 32), float("nan"), device=dev)
    L.call("mas_tc_probe", A, B, D, a_src, b_layout, 0, 0, 0)
    ref = A.double() @ B.double().t()
    err = float((D.double() - ref).norm() / ref.norm())
    print(f"probe a_src={a_src} b_layout={b_layout}: rel err {err:.3e} D[0,:4]={D[0,:4].tolist()} ref={ref[0,:4].tolist()}")
    if b_layout == 0:
        assert err < 2e-3, (a_src, b_layout, err)   # K-major B; A from smem (SS) and from tensor memory (TS) both work
    else:
        # MN-major TF32 shared-memory operands return zeros on this part (any LBO/SBO order, any swizzle): the
        # production kernels therefore keep every smem operand K-major (see wgrad_tc's transposed halo copies)
        assert err > 0.5


@pytest.mark.parametrize("b_layout", [0, 1, 2])
def test_reveal_b_addressing(b_layout):
    """A = selector (row m picks k = m % 8), B region = its own word index: D[k][n] is the shared-memory word the
    tensor core reads for element (n, k) of B under the given descriptor convention."""
    from mas_b200 import _lib as L
    dev = torch.device("cuda:0")
    A = torch.zeros(128, 8)
    for m in range(128):
        A[m, m % 8] = 1.0
    D = torch.full((128, 32), float("nan"), device=dev)
    L.call("mas_tc_probe", A.to(dev), torch.zeros(32, 8, device=dev), D, 0, 10 + b_layout, 0, 0, 0)
    off = D[:8].t().cpu().long()          # [n][k]
    print(f"REVEAL b_layout={b_layout}")
    for n in list(range(0, 10)) + [16, 31]:
        print("  n=%2d:" % n, off[n].tolist())


def _desc(lbo, sbo, layout_type, base_off=0):
    return ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (base_off << 49) | (layout_type << 61)


def _idesc(n, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | ((1 << 16) if b_mn else 0) | ((n >> 3) << 17) | ((128 >> 4) << 24)


RAW = [  # name, lbo, sbo, layout_type, b_mn, start_off, base_off
    ("K-major none (control)", 512, 128, 0, 0, 0, 0),
    ("MN none lbo160 sbo576", 160, 576, 0, 1, 0, 0),
    ("MN sw128 lbo1024 sbo1024", 1024, 1024, 2, 1, 0, 0),
    ("MN sw128 lbo4096 sbo1024", 4096, 1024, 2, 1, 0, 0),
    ("MN sw128 start+128", 4096, 1024, 2, 1, 128, 0),
    ("MN sw128 start+128 base1", 4096, 1024, 2, 1, 128, 1),
    ("MN sw128 start+256", 4096, 1024, 2, 1, 256, 0),
    ("MN sw64 lbo512 sbo512", 512, 512, 4, 1, 0, 0),
    ("MN sw32 lbo256 sbo256", 256, 256, 6, 1, 0, 0),
    ("K sw128 sbo1024", 16, 1024, 2, 0, 0, 0),
    ("K sw128 start+128", 16, 1024, 2, 0, 128, 0),
    ("K sw128 start+128 base1", 16, 1024, 2, 0, 128, 1),
    ("K sw128 start+32 (k advance)", 16, 1024, 2, 0, 32, 0),
]


@pytest.mark.parametrize("case", RAW, ids=[c[0] for c in RAW])
def test_reveal_raw(case):
    from mas_b200 import _lib as L
    name, lbo, sbo, lt, b_mn, off, base = case
    dev = torch.device("cuda:0")
    A = torch.zeros(128, 8)
    for m in range(128):
        A[m, m % 8] = 1.0
    D = torch.full((128, 32), float("nan"), device=dev)
    L.call("mas_tc_probe", A.to(dev), torch.zeros(32, 8, device=dev), D, 0, 99, _desc(lbo, sbo, lt, base), _idesc(32, b_mn), off)
    o = D[:8].t().cpu().long()
    print(f"RAW {name}")
    for n in list(range(0, 9)) + [12, 16, 31]:
        print("  n=%2d:" % n, o[n].tolist())
