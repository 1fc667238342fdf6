#!/bin/bash
# A/B on the GPU box: selected tests (PYTEST_K), then the bench (no CPU baseline) under each listed environment.
# usage: PYTEST_K=expr gpu_ab_quick.sh TAG "ENV1=.. ENV2=.." "ENV3=.." ...   ("" = defaults)
TAG=$1; shift
mkdir -p gpurun_out/$TAG
python -m pytest tests -m gpu -q -x -k "${PYTEST_K:-kernel_variants}" 2>&1 | tail -${TAILN:-8}
i=0
for v in "$@"; do
  env $v python bench.py --steps 10 --warmup 3 --no-secondary ${BENCH_ARGS---no-cpu-baseline} > gpurun_out/$TAG/bench_$i.json 2> gpurun_out/$TAG/bench_$i.err
  echo "variant $i [$v] exit $?"; tail -2 gpurun_out/$TAG/bench_$i.err
  python -c "
import json; d=json.load(open('gpurun_out/$TAG/bench_$i.json')); print(round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],3), 'ms', {k: round(v,3) for k,v in d['config']['phase_ms_per_step'].items()}, 'e2e', round(d['e2e']['value']/1e9,2), 'parity', d['parity']['max_rel_err'] if d.get('parity') else None)"
  i=$((i+1))
done
