mkdir -p gpurun_out/r2w
python -m pytest tests -m gpu -q -x 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/r2w/bench.json 2> gpurun_out/r2w/bench.err; echo "bench exit $?"; tail -2 gpurun_out/r2w/bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2w/bench.json')); print(round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],3), 'ms', {k: round(v,3) for k,v in d['config']['phase_ms_per_step'].items()}, 'e2e', round(d['e2e']['value']/1e9,2), d['parity']['max_rel_err'], d['roofline']['frac'], d['roofline']['whole_step']['frac'])"
python tools/probe_c3shape.py default TSKB_COLS_VARIANT=d default TSKB_COLS_VARIANT=d > gpurun_out/r2w/c3shape.json 2> gpurun_out/r2w/c3shape.err; tail -3 gpurun_out/r2w/c3shape.err; python -c "
import json; d=json.load(open('gpurun_out/r2w/c3shape.json'))
for v,o in d.items():
    print(v)
    for k,x in o.items(): print('  ',k,x)
"
