#!/usr/bin/env bash
# Last pass of the round on HEAD: GPU suite, smoke, the default bench line (headline + transformer block with kernel rooflines).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 90 > $O/r2h_tests.log 2>&1; echo "suite rc=$?"; tail -3 $O/r2h_tests.log | cut -c1-200
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2h_smoke.log | cut -c1-200
timeout 300 python bench.py > $O/r2h_bench.json 2> $O/r2h_bench.err; echo "bench rc=$?"; tail -3 $O/r2h_bench.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("$O/r2h_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","steps","warmup")}, "e2e",d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["clocks"])
print("transformer", json.dumps(d.get("transformer"))[:1800]); print("cpu", d.get("cpu_baseline"))
PY
