#!/usr/bin/env bash
# 2-GPU sanity of the current kernels under DistributedDataParallel: gradient check against a single-process full batch, then
# the bench line at N = 2 (both arms as the driver launches them).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/ddp_check.py 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/ddp_bench_n$N.json
python - <<PY
import json
d=json.loads(open("$O/ddp_bench_n$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches")}, "e2e", d["e2e"]["value"], d["clocks"])
PY
