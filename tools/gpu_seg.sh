#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | grep -B30 "short test summary" | grep "^E \|Error\|^tests.*Error\|assert" | head -40
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
timeout 600 python bench.py --workload vqseg --steps 3 --warmup 2 --profile > $O/seg_bench.json 2> $O/seg_prof.log
tail -1 $O/seg_bench.json | cut -c1-900
head -16 $O/seg_prof.log
