#!/usr/bin/env bash
# Round-2 profile captures (ncu): VQ filter kernel, AttnBlock kernels, whole-step launch list.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== ncu VQ (filter + resolve + exact kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vq_filter_tc|vq_resolve|vq_forward_kernel" -c 4 -o $O/r02_vq python tools/prof_kernels.py vq > $O/r02_vq.log 2>&1; tail -2 $O/r02_vq.log
echo "== ncu AttnBlock kernels (forward + backward of one block)"
timeout 600 ncu --set full --clock-control none -k regex:"gemm3_tc|shift_gemm_tc|softmax|wgrad_tc" -c 24 -o $O/r02_attn python tools/prof_kernels.py attn > $O/r02_attn.log 2>&1; tail -2 $O/r02_attn.log
echo "== launch list of the bench step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02_launches.csv python bench.py --no-cpu-baseline --steps 1 --warmup 1 > $O/r02_launch_bench.log 2>&1; tail -1 $O/r02_launch_bench.log | cut -c1-200
