"""Seeded parameter filling shared by oracle/make_golden.py (run on the REAL reference) and the GPU parity tests (run on
the drop-in modules): both sides fill the same-named parameters from the same CPU generator, so fixtures for wide blocks
need not store the weights (a 512->512 ResnetBlock is 19 MB of them). TEST INFRASTRUCTURE - never imported by the product.
"""
import torch


def fill_seeded(module, seed):
    """Fill every parameter of `module` (state_dict order) from a CPU generator: conv/linear weights ~ N(0, 1/fan_in),
    norm weights 1 + 0.1 N(0,1), biases 0.1 N(0,1). Returns {name: (sum, abs-sum)} in fp64 as a drift check."""
    g = torch.Generator().manual_seed(seed)
    checks = {}
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.dim() >= 2:
                fan_in = p[0].numel()
                v = torch.randn(p.shape, generator=g) / fan_in ** 0.5
            elif "norm" in name and name.endswith("weight"):
                v = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
            else:
                v = 0.1 * torch.randn(p.shape, generator=g)
            p.copy_(v.to(p.device))
            checks[name] = (float(v.double().sum()), float(v.double().abs().sum()))
    return checks


def seeded_input(shape, seed, scale=1.0, shift=0.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale + shift


def sample(t, n=512):
    """Strided sample of a tensor (stride coprime to the usual power-of-two tile sizes) for compact fixtures."""
    f = t.detach().reshape(-1)
    stride = max(1, f.numel() // n)
    if stride > 1 and stride % 2 == 0:
        stride += 1
    return f[::stride].clone(), stride


def assert_same_fill(got, want):
    """fill_seeded check sums agree (same names in the same order; fp64 sums up to summation-order noise)."""
    assert list(got) == list(want), (list(got), list(want))
    for k in got:
        for a, b in zip(got[k], want[k]):
            assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (k, a, b)
