#!/usr/bin/env bash
# TMA-fed fp16 row GEMM behind the transformer's Linear layers: its own test first (own process), the whole GPU suite, then the
# training-step tool with per-entry profile.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 90 python -m pytest tests/test_gpu_transformer.py -q -p no:cacheprovider -k "tma_fed" > $O/r2d_test_gemm.log 2>&1; echo "gemm test rc=$?"; tail -25 $O/r2d_test_gemm.log
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 90 > $O/r2d_tests.log 2>&1; echo "suite rc=$?"; tail -12 $O/r2d_tests.log
timeout 100 python tools/bench_transformer.py --batch 8 --steps 3 --warmup 2 --profile > $O/r2d_tf.json 2> $O/r2d_tf.err; echo "tf rc=$?"; cat $O/r2d_tf.json; grep -v Warning $O/r2d_tf.err | tail -22
