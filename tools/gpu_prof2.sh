#!/bin/bash
# ncu --set full of kernels matching REGEX on the bench workload with optional env. usage: gpu_prof2.sh TAG REGEX SKIP COUNT [ENV...]
TAG=$1; RX=$2; SKIP=$3; CNT=$4; shift 4
mkdir -p gpurun_out/$TAG
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c $CNT \
    -o gpurun_out/$TAG/prof_$RX -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/$TAG/ncu_$RX.log 2>&1
echo "ncu exit $?"
