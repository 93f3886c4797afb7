#!/usr/bin/env bash
# Final round-2 evidence pass: parity suite, smoke, bench (vqimg + vqseg), VQ ncu capture, reference arm.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
TAG=${1:-final}
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 6 --warmup 3 --profile > $O/${TAG}_bench.json 2> $O/${TAG}_prof.log
timeout 600 python bench.py --workload vqseg --steps 4 --warmup 3 > $O/${TAG}_seg.json 2> $O/${TAG}_seg.err
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e",d["e2e"]["value"], "roofline",d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["clocks"], d.get("cpu_baseline"))
v=d["vq"]; print({k:v[k] for k in v if k not in ("sweep","kernel","bound","all_pairs_ffma_kernel")}); print(d["attn"])
s=json.loads(open("$O/${TAG}_seg.json").read().strip().splitlines()[-1]); print("vqseg", s["value"], s["ms_per_step"], s["e2e"])
PY
# ncu --set full of the production convolution kernels on the dominant layer, and the launch list of one step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"shift_gemm_t16|wgrad_t16" -c 6 -f -o $O/r02_conv4 \
  python tools/micro_conv.py one tma_plain tma_stats wgrad16 > $O/r02_conv4.log 2>&1; tail -1 $O/r02_conv4.log
if [ "${2:-}" = "launches" ]; then   # ~8 GPU-minutes: every launch of four steps serialised under ncu
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/r02_launches2.csv \
  python bench.py --no-cpu-baseline --steps 1 --warmup 1 --step-only > $O/r02_launch_bench2.log 2>&1; tail -1 $O/r02_launch_bench2.log | cut -c1-200
fi
