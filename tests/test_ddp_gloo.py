"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path —
(1) the SyncBatchNorm recipe used by models/vqvae.py::QuantBatchNorm (all-reduce of per-rank [sum, sumsq], then
    mean / biased var) reproduces full-batch statistics;
(2) the DDP identity the GPU check (tests/ddp_check.py) relies on: the average over ranks of per-rank gradients equals the
    single-process gradient on the concatenated batch (exercised with the CPU oracle, BN in eval mode)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def _worker(rank, world, port, out):
    sys.path[:0] = [ROOT]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import vqgan_oracle as O
    g = torch.load(os.path.join(GOLDEN, "vqbase_tiny.pt"), weights_only=False)
    X = torch.rand(world * 2, 3, 16, 16, generator=torch.Generator().manual_seed(3))
    x = X[rank * 2:(rank + 1) * 2]
    # (1) BN statistics recipe
    h = torch.randn(world * 2, 32, 4, 4, generator=torch.Generator().manual_seed(9))
    hl = h[rank * 2:(rank + 1) * 2].permute(0, 2, 3, 1).reshape(-1, 32).double()
    stats = torch.cat([hl.sum(0), (hl * hl).sum(0)])
    dist.all_reduce(stats)
    cnt = hl.shape[0] * world
    mean, var = stats[:32] / cnt, stats[32:] / cnt - (stats[:32] / cnt) ** 2
    full = h.permute(0, 2, 3, 1).reshape(-1, 32).double()
    ok1 = torch.allclose(mean, full.mean(0), atol=1e-12) and torch.allclose(var, full.var(0, unbiased=False), atol=1e-12)
    # (2) gradient-averaging identity
    sd = {k: v.clone().requires_grad_(k in g["grads"]) for k, v in g["state_dict"].items()}
    dec, diff, _ = O.vqbase_forward(sd, g["ddconfig"], x, training=False)
    O.proxy_loss(x, dec, diff).backward()
    grads = {k: v.grad.clone() for k, v in sd.items() if v.grad is not None}
    for v in grads.values():
        dist.all_reduce(v)
        v /= world
    ok2 = True
    if rank == 0:
        sd2 = {k: v.clone().requires_grad_(k in g["grads"]) for k, v in g["state_dict"].items()}
        dec, diff, _ = O.vqbase_forward(sd2, g["ddconfig"], X, training=False)
        O.proxy_loss(X, dec, diff).backward()
        for k, v in grads.items():
            ref = sd2[k].grad
            den = max(float(ref.norm()), 1e-5 * ref.numel() ** 0.5)   # ~0 gradients compare on an absolute scale
            if float((v - ref).norm()) > 1e-3 * den:
                ok2 = False
    out.put((rank, bool(ok1), bool(ok2)))
    dist.destroy_process_group()


def test_world2_gloo_host_logic():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29641, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] and r[2] for r in res), res
