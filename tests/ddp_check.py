"""Multi-GPU check (run under torchrun on the GPU box, world_size >= 2):
DDP over the drop-in VQBASE — gradients after the NCCL all-reduce equal the single-process gradients on the concatenated
batch (the SyncBatchNorm statistics are all-reduced inside the forward, so the two are directly comparable).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/ddp_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "make-a-scene_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from models import VQBASE
    dd = dict(z_channels=64, in_channels=3, out_channels=3, channels=[128, 128, 256], num_res_blocks=1, resolution=32,
              attn_resolutions=[16], dropout=0.0)

    def build():
        torch.manual_seed(0)
        m = VQBASE(dd, 512, 64, 10, 100)
        with torch.no_grad():
            m.quantize.embedding.weight.normal_()
        m.quantize.q_counter = 10 ** 6
        return m.train().to(dev)

    per = 2
    X = torch.rand(world * per, 3, 32, 32, generator=torch.Generator().manual_seed(7)).to(dev)
    m = build()
    ddp = torch.nn.parallel.DistributedDataParallel(m, device_ids=[local])
    x = X[rank * per:(rank + 1) * per]
    dec, diff = ddp(x)
    ((x - dec).abs().mean() + diff).backward()
    ok = True
    if rank == 0:
        ref = build()                       # single process, full batch, no process-group use in BN: emulate by one rank
        from mas_b200 import ops
        # full-batch statistics without communication: temporarily hide the process group from the BN function
        saved = ops.dist.is_initialized
        ops.dist.is_initialized = lambda: False
        try:
            dec2, diff2 = ref(X)
            ((X - dec2).abs().mean() + diff2).backward()
        finally:
            ops.dist.is_initialized = saved
        worst = 0.0
        for (k, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
            den = max(float(q.grad.norm()), 1e-5 * q.grad.numel() ** 0.5)
            worst = max(worst, float((p.grad - q.grad).norm()) / den)
        print(f"ddp_check world={world}: worst relative gradient difference vs single-process full batch = {worst:.3e}")
        ok = worst < 5e-3
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
