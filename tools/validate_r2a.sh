#!/usr/bin/env bash
# Round-2 first GPU pass: validate the staged kernels (tc3, then CTA pairs last: it may hang), baseline profile.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== 0. baseline suite"
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== 1. tc3 vs fp64"
MAS_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_staged.py -m gpu -q -k tc3 2>&1 | tail -15
echo "== 2. AttnBlock on tc3: parity suite, bench"
MAS_ATTN_TC3=1 timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
MAS_ATTN_TC3=1 timeout 200 python bench.py --no-cpu-baseline --steps 4 --warmup 3 --profile 2> $O/r2a_prof_tc3.log | tail -1 | cut -c1-400
echo "== 3. baseline bench + per-entry profile"
timeout 200 python bench.py --no-cpu-baseline --steps 4 --warmup 3 --profile 2> $O/r2a_prof_base.log | tail -1 | cut -c1-400
echo "== 4. CTA-pair conv: dedicated shapes"
MAS_EXPERIMENTAL=1 MAS_CONV_2CTA=1 timeout 120 python -m pytest tests/test_gpu_staged.py -m gpu -q -x -k cta_pair 2>&1 | tail -15
rc=${PIPESTATUS[0]}
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
if [ "$rc" = "0" ]; then
  echo "== 5. CTA-pair conv: whole suite + bench"
  MAS_EXPERIMENTAL=1 MAS_CONV_2CTA=1 timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
  MAS_CONV_2CTA=1 timeout 200 python bench.py --no-cpu-baseline --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-400
fi
