#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -12
timeout 400 python bench.py --no-cpu-baseline --steps 4 --warmup 3 --profile > $O/q_bench.json 2> $O/q_prof.log
python - <<PY
import json
d=json.loads(open("$O/q_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["clocks"])
PY
python - <<'PY'
import re,collections
agg=collections.defaultdict(lambda:[0,0.0])
for ln in open('gpurun_out/q_prof.log'):
    m=re.match(r'\s+(\S+)\s+n=\s*(\d+)\s+([\d.]+) ms',ln)
    if not m: continue
    name=m.group(1).split('|')[0]
    agg[name][0]+=int(m.group(2)); agg[name][1]+=float(m.group(3))
tot=sum(v[1] for v in agg.values())
print("total",round(tot,2))
for k,v in sorted(agg.items(),key=lambda kv:-kv[1][1])[:10]:
    print("   %-34s n=%4d %8.2f ms %5.1f%%"%(k,v[0],v[1],100*v[1]/tot))
PY
