#!/bin/bash
# usage: tools/prof_full.sh <tag> <workload> <kernel-regex> [launch-skip]   -> gpurun_out/<tag>_full.ncu-rep of one launch
tag=$1; wl=$2; k=$3; skip=${4:-3}
mkdir -p gpurun_out
VRS_NO_GRAPH=1 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip $skip -c 1 -f -o gpurun_out/${tag}_full \
  python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_stdout.log 2>&1
ls -la gpurun_out/${tag}_full.ncu-rep
