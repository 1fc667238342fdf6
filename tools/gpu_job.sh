mkdir -p gpurun_out/r2z0
python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2z0/bench.json 2> gpurun_out/r2z0/bench.err; echo "bench exit $?"; tail -2 gpurun_out/r2z0/bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2z0/bench.json')); print(round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],3), 'ms', {k: round(v,3) for k,v in d['config']['phase_ms_per_step'].items()}, 'e2e', round(d['e2e']['value']/1e9,2))"
timeout 900 python bench.py --config c3 --steps 4 --warmup 1 > gpurun_out/r2z0/c3_n1.json 2> gpurun_out/r2z0/c3_n1.err; echo "c3 N=1 exit $?"
head -c 2500 gpurun_out/r2z0/c3_n1.json; echo; tail -3 gpurun_out/r2z0/c3_n1.err
