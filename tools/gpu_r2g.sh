#!/usr/bin/env bash
# Causal block skipping in the attention GEMMs / softmax backward: transformer + GEMM tests, then the training-step tool.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 150 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_gemm3.py -q -p no:cacheprovider --timeout 60 > $O/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -6 $O/r2g_tests.log | cut -c1-300
timeout 100 python tools/bench_transformer.py --batch 8 --steps 3 --warmup 2 --profile > $O/r2g_tf.json 2> $O/r2g_tf.err; echo "tf rc=$?"; cut -c1-330 $O/r2g_tf.json; grep -v "Warning\|Consider\|unit" $O/r2g_tf.err | head -9
