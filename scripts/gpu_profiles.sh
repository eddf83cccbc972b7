#!/bin/bash
# Run under gpurun: launch list of bench.py + ncu --set full captures of the two dominant kernels.
# usage: bash scripts/gpu_profiles.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"head_|hist_" -c 60 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 8 --warmup 3 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_kernel -s 3 -c 1 -o gpurun_out/head_fwdbwd_${TAG} \
    python scripts/ncu_target.py head 8 0 1 > gpurun_out/ncu_head_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hist_u8 -s 2 -c 1 -o gpurun_out/hist_${TAG} \
    python scripts/ncu_target.py hist 64 1 > gpurun_out/ncu_hist_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
