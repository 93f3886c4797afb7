#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "codebook" 2>&1 | grep -v Warning | tail -30
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -8
timeout 400 python bench.py --no-cpu-baseline --steps 4 --warmup 3 > $O/vq_bench.json 2> $O/vq_bench.err
python - <<PY
import json
d=json.loads(open("$O/vq_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"])
v=d["vq"]; print({k:v[k] for k in v if k not in ("sweep",)}); print([(p["batch"],p["ms"],p["tflop_per_s"],p["gb_per_s"]) for p in v.get("sweep",[])])
PY
tail -5 $O/vq_bench.err
