#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== f16 (incl. wgrad): whole suite"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -40
echo "== tf32: whole suite"
MAS_CONV_OPERANDS=tf32 timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -8
echo "== micro"
timeout 600 python tools/micro_conv.py 2>&1 | tail -50
echo "== bench f16"
timeout 400 python bench.py --no-cpu-baseline --steps 4 --warmup 3 --profile > $O/f16b_bench.json 2> $O/f16b_prof.log
python - <<PY
import json
d=json.loads(open("$O/f16b_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["clocks"])
PY
head -24 $O/f16b_prof.log
echo "== ncu f16 plain / pro / wgrad"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"shift_gemm_tc|wgrad_tc" -c 6 -o $O/f16b_conv python tools/micro_conv.py one f16_plain f16_pro f16_wgrad > $O/f16b_ncu.log 2>&1
tail -3 $O/f16b_ncu.log
