#!/bin/bash
# Multi-GPU visit: the genome-sharded bench line, the sharded-vs-one-GPU check of every sharded path, C3.
# Usage: gpurun --gpus N --timeout 1800 -- 'bash tools/gpu_multi.sh r2m2 N [c3]'
TAG=${1:-r2m}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/smi.txt 2>&1; (nproc; free -g | head -2) >> $OUT/smi.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench N=$N exit $?"
head -c 2500 $OUT/bench_n$N.json; echo
timeout 900 $TR tools/shard_check.py > $OUT/shard_check_n$N.json 2> $OUT/shard_check_n$N.err; echo "shard_check exit $?"
head -c 3000 $OUT/shard_check_n$N.json; echo; tail -3 $OUT/shard_check_n$N.err
if [ "$3" == "c3" ]; then
  timeout 600 $TR bench.py --config c3 --small --steps 3 --warmup 1 > $OUT/c3_small_n$N.json 2> $OUT/c3_small_n$N.err; echo "c3 small exit $?"
  head -c 1500 $OUT/c3_small_n$N.json; echo; tail -3 $OUT/c3_small_n$N.err
  timeout 1500 $TR bench.py --config c3 --steps 3 --warmup 1 > $OUT/c3_n$N.json 2> $OUT/c3_n$N.err; echo "c3 exit $?"
  head -c 3000 $OUT/c3_n$N.json; echo; tail -5 $OUT/c3_n$N.err
fi
ls -la $OUT
