#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -8
timeout 600 python tools/micro_conv.py 2>&1 | head -22
timeout 400 python bench.py --no-cpu-baseline --steps 4 --warmup 3 --profile > $O/mix_bench.json 2> $O/mix_prof.log
python - <<PY
import json
d=json.loads(open("$O/mix_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["clocks"])
v=d["vq"]; print({k:v[k] for k in v if k not in ("sweep","kernel","bound")}); print([(p["batch"],p["ms"],p["tflop_per_s"],p["gb_per_s"]) for p in v.get("sweep",[])])
PY
head -14 $O/mix_prof.log
