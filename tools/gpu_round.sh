#!/bin/bash
# One GPU-box visit: parity suite, bench (both arms), ncu launch list, ncu full capture.
# Usage: gpurun --timeout 1800 -- 'bash tools/gpu_round.sh r01a'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json
if [ -z "$SKIP_REF" ]; then
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref exit $?"
cat $OUT/bench_ref.json
fi
timeout 300 python tools/probe_relvec.py > $OUT/probe_relvec.json 2> $OUT/probe_relvec.err; echo "relvec probe exit $?"
cat $OUT/probe_relvec.json
PROBE_QUICK=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'(k_relvec|k_sweep|k_init_weights)' -c 100 --csv --log-file $OUT/relvec_launches.csv python tools/probe_relvec.py > $OUT/ncu_relvec.log 2>&1
echo "ncu relvec exit $?"
KREGEX='regex:(k_set_weights|k_sweep|k_branch_summary|k_window|k_site_summary|DeviceScan)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 600 \
    --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 2 -c 2 \
    -o $OUT/prof_sweep -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_prop.log 2>&1
echo "ncu sweep exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_branch_summary -s 2 -c 2 \
    -o $OUT/prof_summary -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_sum.log 2>&1
echo "ncu summary exit $?"
ls -la $OUT
