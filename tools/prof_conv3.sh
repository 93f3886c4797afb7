#!/usr/bin/env bash
# ncu --set full of the production convolution kernels on the dominant layer (conv3x3 128->128 @256^2, batch 32):
# TMA-fed fprop (plain / residual + statistics epilogue), TMA-fed data gradient, weight gradient fed by fp16 shadows.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"shift_gemm_t16|wgrad_tc" -c 6 -o $O/r02_conv3 \
  python tools/micro_conv.py one tma_plain tma_stats wgrad16 > $O/r02_conv3.log 2>&1
tail -2 $O/r02_conv3.log
