#!/bin/bash
# ncu --set full of one kernel regex on the bench workload. usage: gpu_prof.sh TAG REGEX [skip] [count]
TAG=$1; RX=$2; SKIP=${3:-2}; CNT=${4:-1}
mkdir -p gpurun_out/$TAG
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c $CNT \
    -o gpurun_out/$TAG/prof_$RX -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/$TAG/ncu_$RX.log 2>&1
echo "ncu exit $?"
