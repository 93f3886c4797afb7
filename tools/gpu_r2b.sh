#!/usr/bin/env bash
# Round-2 late pass (few GPU-minutes left): parity suite incl. the new transformer entries, headline bench line,
# transformer training-step and sampling throughput.  Every leg under its own timeout; outputs under gpurun_out/.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
TAG=${1:-r2b}
timeout 240 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 90 > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -25 $O/${TAG}_tests.log
timeout 200 python bench.py --steps 6 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
timeout 100 python tools/bench_transformer.py --batch 8 --steps 3 --warmup 2 > $O/${TAG}_tf.json 2> $O/${TAG}_tf.err; echo "tf rc=$?"; cat $O/${TAG}_tf.json; tail -3 $O/${TAG}_tf.err
timeout 100 python tools/bench_sampling.py --graphs --reps 2 > $O/${TAG}_sample.json 2> $O/${TAG}_sample.err; echo "sample rc=$?"; cat $O/${TAG}_sample.json; tail -3 $O/${TAG}_sample.err
python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e",d["e2e"]["value"], "roofline",d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["clocks"], d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e); print(open("$O/${TAG}_bench.err").read()[-1500:])
PY
