#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_path_blocks or img_config" 2>&1 | grep -v Warning | tail -25
timeout 120 python - <<'PY'
import sys
sys.path[:0]=['.','make-a-scene_b200']
import torch, bench
dev=torch.device('cuda:0')
print(bench.attn_metric(dev, bench.peaks()))
import os
PY
MAS_ATTN_FUSED=0 timeout 120 python - <<'PY'
import sys
sys.path[:0]=['.','make-a-scene_b200']
import torch, bench
dev=torch.device('cuda:0')
print("unfused", bench.attn_metric(dev, bench.peaks()))
PY
