#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== tf32: tensor-path block fixtures (full output)"
MAS_CONV_OPERANDS=tf32 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tensor_path_blocks or img_config" 2>&1 | grep -v Warning | tail -60
echo "== f16: whole suite"
MAS_CONV_OPERANDS=f16 timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -80
echo "== bench tf32 / f16"
for fmt in tf32 f16; do
MAS_CONV_OPERANDS=$fmt timeout 400 python bench.py --no-cpu-baseline --steps 4 --warmup 3 --profile > $O/f16_${fmt}_bench.json 2> $O/f16_${fmt}_prof.log
python - <<PY
import json
d=json.loads(open("$O/f16_${fmt}_bench.json").read().strip().splitlines()[-1])
print("$fmt", {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["clocks"])
PY
head -14 $O/f16_${fmt}_prof.log
done
