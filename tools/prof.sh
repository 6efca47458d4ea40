#!/bin/bash
# usage: tools/prof.sh <tag> <workload> [kernel-regex for one --set full capture] [launch-skip]
# Writes gpurun_out/<tag>_launches.csv (per-launch metrics of frames 3-4, --clock-control none, kernels launched one by one:
# VRS_NO_GRAPH=1 VRS_PIPELINE=0) and, when a kernel regex is given, gpurun_out/<tag>_<regex>_full.ncu-rep of one launch of that
# kernel.  Numbers printed under ncu are not bench values: cold caches, serialised launches — compare SHARES.
tag=$1; wl=$2; k=$3; skip=${4:-3}
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
export VRS_NO_GRAPH=1 VRS_PIPELINE=0
timeout 600 ncu --metrics $M --clock-control none --launch-skip 36 -c 26 --csv --log-file gpurun_out/${tag}_launches.csv \
  python tools/ncu_frames.py $wl 6 > gpurun_out/${tag}_ncu_stdout.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv
for kk in $(echo $k | tr ',' ' '); do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:"$kk" --launch-skip $skip -c 1 -f -o gpurun_out/${tag}_${kk}_full \
    python tools/ncu_frames.py $wl 5 >> gpurun_out/${tag}_ncu_stdout.log 2>&1
  ncu -i gpurun_out/${tag}_${kk}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_${kk}_full_raw.csv 2>/dev/null
done
