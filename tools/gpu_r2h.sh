#!/usr/bin/env bash
# Last seconds of the round's GPU budget: the transformer block of bench.py (the only code not yet run on a GPU) in isolation.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 120 python - > $O/r2h_tf_block.json 2> $O/r2h_tf_block.err <<PY
import json, torch, bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
print(json.dumps(bench.transformer_metric(dev, bench.peaks())))
PY
echo "rc=$?"; cat $O/r2h_tf_block.json | cut -c1-2500; tail -3 $O/r2h_tf_block.err | cut -c1-300
