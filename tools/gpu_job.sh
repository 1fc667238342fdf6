mkdir -p gpurun_out/r2x
for v in "TSKB_RUN_COPIES=128" "TSKB_RUN_COPIES=1024" "TSKB_RUN_COPIES=512" "TSKB_RUN_COPIES=32" "TSKB_SUM_VARIANT=lane"; do
env $v python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2x/bench.json 2> gpurun_out/r2x/bench.err; echo "bench [$v] exit $?"; tail -2 gpurun_out/r2x/bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2x/bench.json')); print(round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],3), 'ms', {k: round(v,3) for k,v in d['config']['phase_ms_per_step'].items()}, 'e2e', round(d['e2e']['value']/1e9,2))"
done
