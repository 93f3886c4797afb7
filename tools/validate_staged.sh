#!/usr/bin/env bash
# Validation recipe for the staged kernels.  Run on a B200 box, e.g.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/validate_staged.sh > gpurun_out/staged.log 2>&1; tail -40 gpurun_out/staged.log'
# Every step is wrapped in `timeout`: a wrong barrier protocol in the CTA-pair kernel hangs instead of failing.
set -u
cd "$(dirname "$0")/.."

echo "== 1. 3xTF32 batched GEMM (contract_tc3.cu) against fp64"
MAS_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_staged.py -m gpu -q -k tc3 2>&1 | tail -5

echo "== 2. cta_group::2 convolution (contract_tc2.cu): dedicated shapes first, then the whole parity suite on it"
MAS_EXPERIMENTAL=1 MAS_CONV_2CTA=1 timeout 120 python -m pytest tests/test_gpu_staged.py -m gpu -q -k cta_pair 2>&1 | tail -5
MAS_EXPERIMENTAL=1 MAS_CONV_2CTA=1 timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -5

echo "== 2b. AttnBlock token contractions on the 3xTF32 kernel (MAS_ATTN_TC3=1): parity suite, then step time"
MAS_ATTN_TC3=1 timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
MAS_ATTN_TC3=1 timeout 200 python bench.py --no-cpu-baseline --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-260

echo "== 3. step time with / without the CTA-pair kernel (same box)"
timeout 200 python bench.py --no-cpu-baseline --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-260
MAS_CONV_2CTA=1 timeout 200 python bench.py --no-cpu-baseline --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-260
