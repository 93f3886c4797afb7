#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "codebook or vqbase_tiny or kmeans" 2>&1 | grep -v Warning | tail -12
timeout 200 python - <<'PY'
import sys, json
sys.path[:0]=['.','make-a-scene_b200']
import torch, bench
dev=torch.device('cuda:0')
v=bench.vq_metric(dev, bench.peaks())
print({k:v[k] for k in v if k not in ("sweep","kernel","bound","all_pairs_ffma_kernel")})
print([(p["batch"],p["ms"],p["tflop_per_s"],p["gb_per_s"],p["tensor_frac_3pass"]) for p in v["sweep"]])
PY
