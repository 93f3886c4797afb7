#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -12
timeout 600 python tools/micro_conv.py 2>&1 | head -12
timeout 400 python bench.py --no-cpu-baseline --steps 4 --warmup 3 --profile > $O/f16c_bench.json 2> $O/f16c_prof.log
python - <<PY
import json
d=json.loads(open("$O/f16c_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["clocks"])
PY
head -16 $O/f16c_prof.log
