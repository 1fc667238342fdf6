#!/bin/bash
# A/B on the GPU box: parity suite, then the bench (no CPU baseline) under each listed environment.
# usage: gpu_ab.sh TAG "ENV1=.. ENV2=.." "ENV3=.." ...   (first variant "" = defaults)
TAG=$1; shift
mkdir -p gpurun_out/$TAG
python -m pytest tests -m gpu -q 2>&1 | tail -${TAILN:-25}
i=0
for v in "$@"; do
  env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/$TAG/bench_$i.json 2> gpurun_out/$TAG/bench_$i.err
  echo "variant $i [$v] exit $?"; tail -2 gpurun_out/$TAG/bench_$i.err
  python -c "
import json; d=json.load(open('gpurun_out/$TAG/bench_$i.json')); print(round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],3), 'ms', d['config']['phase_ms_per_step'], 'e2e', round(d['e2e']['value']/1e9,2), 'stage_s', round(d['config']['stage_s'],2))"
  i=$((i+1))
done
