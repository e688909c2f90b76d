#!/bin/bash
# Runs on the GPU box (via gpurun): ncu launch list + ncu --set full of the step kernel (bench.py cfg2 workload).
# usage: scripts/gpu_profile.sh <tag> [kernel-regex] [workload]
TAG=${1:-run}
KREGEX=${2:-env_kernelIfLi16ELi3ELi0ELi8101}
WL=${3:-cfg2}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:env_kernel -s 320 -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --workload $WL --steps 20 --warmup 5 --small-ring --no-cpu-baseline --no-fp64 > gpurun_out/ncu1_${TAG}.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:${KREGEX} -s 310 -c 1 -o gpurun_out/prof_${TAG} python bench.py --workload $WL --steps 10 --warmup 5 --small-ring --no-cpu-baseline --no-fp64 > gpurun_out/ncu2_${TAG}.log 2>&1
tail -2 gpurun_out/ncu2_${TAG}.log
