#!/usr/bin/env bash
# Round-2 closing pass (the recipe behind profiles/r02b_*, r02_transformer_*, r02_ncu_transformer.md): GPU suite, smoke, headline
# bench, transformer training / sampling tools, ncu --set full of the new kernels.  tools/gpu_r2g.sh: the later check of the causal
# block skipping.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 240 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 90 > $O/r2e_tests.log 2>&1; echo "suite rc=$?"; tail -8 $O/r2e_tests.log | cut -c1-300
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2e_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/r2e_smoke.log | cut -c1-300
timeout 200 python bench.py --steps 6 --warmup 3 > $O/r2e_bench.json 2> $O/r2e_bench.err; echo "bench rc=$?"
timeout 100 python tools/bench_transformer.py --batch 8 --steps 3 --warmup 2 --profile > $O/r2e_tf.json 2> $O/r2e_tf.err; echo "tf rc=$?"; cat $O/r2e_tf.json; grep -v Warning $O/r2e_tf.err | tail -20
MAS_TC3_SHALLOW=0 timeout 100 python tools/bench_transformer.py --batch 8 --steps 3 --warmup 2 > $O/r2e_tf_noshallow.json 2> /dev/null; echo "tf(noshallow) rc=$?"; cut -c1-260 $O/r2e_tf_noshallow.json
timeout 100 python tools/bench_sampling.py --graphs --reps 2 > $O/r2e_sample.json 2> $O/r2e_sample.err; echo "sample rc=$?"; cut -c1-330 $O/r2e_sample.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"rows_gemm_t16|rows_wgrad_t16|attn_causal_fwd" -c 16 -f -o $O/r02_tf \
  python tools/bench_transformer.py --layers 1 --batch 8 --steps 1 --warmup 0 > $O/r02_tf_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $O/r02_tf_ncu.log | cut -c1-200
python - <<PY
import json
try:
    d=json.loads(open("$O/r2e_bench.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e",d["e2e"]["value"], "roofline",d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["clocks"], d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e); print(open("$O/r2e_bench.err").read()[-1500:])
PY
