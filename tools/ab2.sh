#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in "$@"; do
  env $v python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/ab_smoke.json 2> gpurun_out/ab_err.log
  env $v python bench.py --workload bunny_4k_full --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/ab_bunny.json 2>> gpurun_out/ab_err.log
  python - "$v" <<'PY'
import json, sys
for n in ["smoke", "bunny"]:
    try:
        d = json.loads(open("gpurun_out/ab_%s.json" % n).read().strip().splitlines()[-1])
        print(sys.argv[1], n, d["value"], d["ms_per_step"], d["pass_ms"], "e2e", d["e2e"]["value"])
    except Exception as e:
        print(sys.argv[1], n, "failed", e, open("gpurun_out/ab_err.log").read()[-2000:])
PY
done
