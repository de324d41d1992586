#!/bin/bash
# tools/launch_list.sh <tag> [bench args]: device time of every kernel launch of a short bench run (ncu, cold-cache and
# serialised: compare shares), gpurun_out/launches_<tag>.csv.  The plain run goes first (the profiling recipe's rule).
TAG=$1; shift
export IRONB_BENCH_MIN_WARMUP_S=0
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-clocks $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "launch list $TAG rc=$?"; wc -l gpurun_out/launches_$TAG.csv
