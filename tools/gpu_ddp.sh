#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/ddp_check.py 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/ddp_trace.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -45
ls -la $O/ddp_trace_w$N.json 2>/dev/null && gzip -f $O/ddp_trace_w$N.json
