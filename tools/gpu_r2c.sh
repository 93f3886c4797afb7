#!/usr/bin/env bash
# Fused causal attention core: its parity test in its own process (a hang costs 90 s), then the training-step tool with
# the fused core and with the GEMM / softmax / GEMM sequence, per-entry profile of each.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 90 python -m pytest tests/test_gpu_transformer.py -q -p no:cacheprovider -k "causal_attention or fused_causal" > $O/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -30 $O/r2c_tests.log
MAS_ATTN_FUSED=0 timeout 100 python tools/bench_transformer.py --batch 8 --steps 3 --warmup 2 --profile > $O/r2c_tf_gemm.json 2> $O/r2c_tf_gemm.err; echo "tf gemm rc=$?"; cat $O/r2c_tf_gemm.json; grep -v Warning $O/r2c_tf_gemm.err | tail -28
timeout 100 python tools/bench_transformer.py --batch 8 --steps 3 --warmup 2 --profile > $O/r2c_tf_fused.json 2> $O/r2c_tf_fused.err; echo "tf fused rc=$?"; cat $O/r2c_tf_fused.json; grep -v Warning $O/r2c_tf_fused.err | tail -28
