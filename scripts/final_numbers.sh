#!/bin/bash
# One pass over everything that goes into profiles/ for the round: bench lines (both arms, all workloads) and the ncu captures.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench.err | grep '^{' > gpurun_out/${TAG}_bench_cfg2.json
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 2>>gpurun_out/${TAG}_bench.err | grep '^{' > gpurun_out/${TAG}_bench_reference.json
for wl in cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-fp64 2>>gpurun_out/${TAG}_bench.err | grep '^{' > gpurun_out/${TAG}_bench_${wl}.json
done
bash scripts/gpu_profile.sh ${TAG}
python - <<PY
import json
for wl in ('cfg2','cfg3','cfg4','cfg5'):
    try:
        d=json.loads(open(f'gpurun_out/${TAG}_bench_{wl}.json').read())
        print(wl, round(d['value']/1e6,2), 'M pipelined;', round(d['serialized']['value']/1e6,2), 'M serialized; e2e', round(d['e2e']['value']/1e6,2), 'M', d['timing']['step_kernel_variant'])
    except Exception as e: print(wl, 'failed', e)
PY
tail -3 gpurun_out/${TAG}_bench.err
