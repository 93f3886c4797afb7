#!/usr/bin/env bash
# Standard GPU pass: parity suite, smoke(), bench (JSON line + per-entry profile). Logs under gpurun_out/<tag>_*.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 400 python bench.py --no-cpu-baseline --steps 4 --warmup 3 --profile > $O/${TAG}_bench.json 2> $O/${TAG}_prof.log
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["clocks"])
v=d["vq"]; print({k:v[k] for k in v if k!="sweep"}); print([(p["batch"],p["ms"],p["tflop_per_s"]) for p in v.get("sweep",[])])
PY
head -12 $O/${TAG}_prof.log
