#!/bin/bash
# One GPU-box visit: parity suite, bench (both arms), A/B of kernel variants, init phases, ncu.
# Usage: gpurun --timeout 1800 -- 'bash tools/gpu_round.sh r2a [quick]'
TAG=${1:-r2}
MODE=${2:-full}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/smi.txt 2>&1
(nproc; free -g; lscpu | head -20; df -h /tmp | tail -1) > $OUT/host.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
head -c 3000 $OUT/bench.json; echo
# A/B of the branch-summary variants (same box, same plan; device-timed, no CPU legs)
[ "$MODE" == "final" ] || for v in "runs_default:" "lane:TSKB_SUM_VARIANT=lane" "c4:TSKB_SUM_VARIANT=c4"; do
    name=${v%%:*}; envs=${v#*:}
    if [ -n "$envs" ] && [[ "$envs" == TSKB_LIB=* ]] && [ ! -f "${envs#TSKB_LIB=}" ]; then continue; fi
    env $envs python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/ab_$name.json 2> $OUT/ab_$name.err
    python - "$OUT/ab_$name.json" "$name" <<'EOF'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f" % d["ms_per_step"], {k: round(v, 4) for k, v in d["config"]["phase_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[2], "failed", e)
EOF
done
if [ -n "$FULLER" ] && [ -x tools/mb/red_types ]; then timeout 120 tools/mb/red_types > $OUT/red_types.txt 2>&1; cat $OUT/red_types.txt; fi
# the other config shapes (C1, C3 shape on the C2 ARG, C5), and the many-column path old vs new
[ "$MODE" == "final" ] || (timeout 600 python tools/probe_configs.py > $OUT/probe_configs.json 2> $OUT/probe_configs.err; echo "probe exit $?")
[ -n "$FULLER" ] && TSKB_COLS_VARIANT=old TSKB_SUM_VARIANT=lane timeout 600 python tools/probe_configs.py > $OUT/probe_configs_oldcols.json 2> $OUT/probe_configs_oldcols.err
[ -n "$FULLER" ] && TSKB_COLS_MIN=2 timeout 600 python tools/probe_configs.py > $OUT/probe_configs_colsmin2.json 2> $OUT/probe_configs_colsmin2.err
python - $OUT <<'EOF2'
import json, sys
for f in ("probe_configs.json", "probe_configs_oldcols.json", "probe_configs_colsmin2.json"):
    try:
        d = json.load(open(sys.argv[1] + "/" + f))
        c3 = d["c3_shape_on_c2_arg_8_sets"]
        print(f, {k: (round(v, 3) if isinstance(v, float) else [round(x, 3) for x in v]) for k, v in c3.items() if "branch" in k})
        print("   c5", d["c5_custom_summary_1e6_windows"])
    except Exception as e:
        print(f, "failed", e)
EOF2
# phases of tskb_treeseq_init on C2 (second init in the process: context and modules already loaded)
[ "$MODE" == "final" ] || TSKB_TIMING=1 python - > $OUT/init_phases.txt 2>&1 <<'EOF'
import time, bench
from tskit_b200.lowlevel import LLTreeSequence
t, W, _ = bench.load_workload("c2")
for i in range(3):
    t0 = time.perf_counter(); ll = LLTreeSequence(t); dt = time.perf_counter() - t0
    print("init %d: %.3f s (engine %.3f s)" % (i, dt, ll.engine_stats()["stage_ms"] / 1e3), flush=True)
    ll.close()
EOF
grep -E "^init|tskb init" $OUT/init_phases.txt | tail -24
if [ "$MODE" == "matrixq" ]; then
timeout 900 python tools/bench_matrix.py > $OUT/bench_matrix_c4.json 2> $OUT/bench_matrix_c4.err; echo "matrix exit $?"; cat $OUT/bench_matrix_c4.json
fi
if [ "$MODE" == "matrix" ]; then
# C4: 20 000 samples x 10^6 biallelic sites; TMA kernel vs the cp.async kernel, then the ncu capture
timeout 900 python tools/bench_matrix.py > $OUT/bench_matrix_c4.json 2> $OUT/bench_matrix_c4.err; echo "matrix exit $?"; cat $OUT/bench_matrix_c4.json
timeout 900 python tools/bench_matrix.py --path cpasync > $OUT/bench_matrix_c4_cpasync.json 2> $OUT/bench_matrix_c4_cpasync.err; cat $OUT/bench_matrix_c4_cpasync.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gram -c 1 -o $OUT/prof_gram -f \
    python tools/bench_matrix.py --samples 16000 --sites 500000 --reps 1 > $OUT/ncu_gram.log 2>&1; echo "ncu gram exit $?"
ncu -i $OUT/prof_gram.ncu-rep --page raw --csv > $OUT/prof_gram_raw.csv 2>/dev/null
python - $OUT/prof_gram_raw.csv <<'EOF2'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
if len(rows) > 2:
    d = dict(zip(rows[0], rows[2]))
    for k in ("Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
              "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):
        print(k, d.get(k))
    for k, v in d.items():
        if "tensor" in k and "pct" in k:
            print(k, v)
EOF2
rm -f $OUT/prof_gram.ncu-rep.tmp
fi
if [ "$MODE" == "c3" ]; then
timeout 1500 python bench.py --config c3 --steps 3 --warmup 1 > $OUT/c3_n1.json 2> $OUT/c3_n1.err; echo "c3 N=1 exit $?"
head -c 3000 $OUT/c3_n1.json; echo; tail -5 $OUT/c3_n1.err
fi
if [ "$MODE" == "full" ] || [ "$MODE" == "final" ]; then
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref exit $?"
cat $OUT/bench_ref.json
[ -n "$FULLER" ] && (timeout 300 python tools/piece_stats.py > $OUT/piece_stats.txt 2>&1; tail -12 $OUT/piece_stats.txt)
KREGEX='regex:(k_set_weights|k_sweep|k_branch_summary|k_runs|k_window|k_site_summary|DeviceScan)'
TSKB_BENCH_BLOCKS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 600 \
    --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/ncu_launch.log 2>&1
echo "ncu launches exit $?"
TSKB_BENCH_BLOCKS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_branch_summary -s 6 -c 2 \
    -o $OUT/prof_summary -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/ncu_sum.log 2>&1
echo "ncu summary exit $?"
TSKB_BENCH_BLOCKS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 6 -c 2 \
    -o $OUT/prof_sweep -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/ncu_prop.log 2>&1
echo "ncu sweep exit $?"
timeout 900 python bench.py --config c3 --steps 6 --warmup 2 > $OUT/c3_n1.json 2> $OUT/c3_n1.err; echo "c3 N=1 exit $?"
head -c 2600 $OUT/c3_n1.json; echo; tail -3 $OUT/c3_n1.err
timeout 300 python tools/probe_c3shape.py default TSKB_COLS_VARIANT=d > $OUT/c3shape.json 2> $OUT/c3shape.err; cat $OUT/c3shape.json
fi
# the .ncu-rep files (60 MB each) do not travel back: their summaries do
if ls $OUT/*.ncu-rep > /dev/null 2>&1; then
  python tools/ncu_summary.py $TAG > /dev/null 2>&1
  mkdir -p $OUT/profiles; cp profiles/${TAG}_* profiles/traffic.json $OUT/profiles/ 2>/dev/null
  for r in $OUT/*.ncu-rep; do python tools/ncu_sass_hot.py $r > $OUT/profiles/${TAG}_$(basename $r .ncu-rep)_sass_hot.txt 2>/dev/null; done
  rm -f $OUT/*.ncu-rep
fi
ls -la $OUT $OUT/profiles 2>/dev/null
