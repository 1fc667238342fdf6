#!/bin/bash
# quick GPU check: timeline of the propagation, parity suite, bench without the CPU baseline
TAG=${1:-quick}
mkdir -p gpurun_out/$TAG
python tools/trace_levels.py c2 | tail -22
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/$TAG/bench.json 2> gpurun_out/$TAG/bench.err
python -c "
import json; d=json.load(open('gpurun_out/$TAG/bench.json')); print(d['value']/1e9, d['ms_per_step'], d['config']['phase_ms_per_step'], d['e2e']['value']/1e9, d['roofline']['whole_sweep'])"
