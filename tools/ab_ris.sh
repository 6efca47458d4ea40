#!/bin/bash
# A/B of the RIS stage (serial per-thread loop vs. the cooperative kernel) and of the small-launch lane limit.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for v in thread coop; do
  VRS_RIS=$v python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/ab_smoke_$v.json 2> gpurun_out/ab_err_$v.log
  VRS_RIS=$v python bench.py --workload bunny_4k_full --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/ab_bunny_$v.json 2>> gpurun_out/ab_err_$v.log
done
for w in 0 8 16 32; do
  echo "min_warps_per_sm $w: $(VRS_MIN_WARPS_PER_SM=$w python tools/band_probe.py 500 620 40)"
done
python - <<'PY'
import json
for n in ["smoke_thread", "smoke_coop", "bunny_thread", "bunny_coop"]:
    try:
        d = json.loads(open("gpurun_out/ab_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["pass_ms"], "e2e", d["e2e"]["value"])
    except Exception as e:
        print(n, "failed", e)
PY
