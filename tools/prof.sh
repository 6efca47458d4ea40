#!/bin/bash
# usage: tools/prof.sh <tag> <workload> [kernel-regex for one --set full capture]
# Writes gpurun_out/<tag>_launches.csv (per-launch metrics, --clock-control none) and, when a kernel regex is given,
# gpurun_out/<tag>_full.ncu-rep (+ raw csv) of one launch of that kernel.  Numbers printed under ncu are not bench values.
tag=$1; wl=$2; k=$3
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
VRS_NO_GRAPH=1 ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_stdout.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv
if [ -n "$k" ]; then
  VRS_NO_GRAPH=1 ncu --set full --import-source on --clock-control none -k regex:"$k" --launch-skip 6 -c 2 -f -o gpurun_out/${tag}_full \
    python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/${tag}_ncu_stdout.log 2>&1
  ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
fi
