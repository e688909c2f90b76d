#!/bin/bash
# Runs on the GPU box (via gpurun): short bench, ncu launch list, ncu --set full of the step kernel.
# usage: scripts/gpu_profile.sh <tag>
TAG=${1:-run}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 500 --warmup 50 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cat gpurun_out/bench_${TAG}.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','ms_per_step_l2_flushed')}, d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['cpu_baseline'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:env_kernel -s 40 -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 40 --warmup 20 --small-ring --no-cpu-baseline > gpurun_out/ncu1_${TAG}.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:env_kernelIfLi16ELi3ELi0 -s 15 -c 1 -o gpurun_out/prof_${TAG} python bench.py --steps 10 --warmup 10 --small-ring --no-cpu-baseline > gpurun_out/ncu2_${TAG}.log 2>&1
tail -1 gpurun_out/ncu2_${TAG}.log
