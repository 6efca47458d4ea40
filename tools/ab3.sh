#!/bin/bash
# usage: tools/ab3.sh <workload> <steps> ENV=val[,ENV2=val2] ...   one bench line per environment (no tests)
wl=$1; steps=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  env ${v//,/ } python bench.py --workload $wl --steps $steps --warmup 5 --no-cpu-baseline > gpurun_out/ab3.json 2> gpurun_out/ab3_err.log
  python - "$v" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/ab3.json").read().strip().splitlines()[-1])
    print(sys.argv[1], d["config"]["name"], d["value"], d["ms_per_step"], d["pass_ms"], "e2e", d["e2e"]["value"])
except Exception as e:
    print(sys.argv[1], "failed", e, open("gpurun_out/ab3_err.log").read()[-2000:])
PY
done
