#!/bin/bash
# Round-end single-GPU evidence: tests, bench lines (ours + reference arm), per-launch lists, one --set full capture.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r01b_pytest_gpu.log 2>&1; tail -3 gpurun_out/r01b_pytest_gpu.log
python bench.py > gpurun_out/r01b_bench_default.json 2> gpurun_out/r01b_bench_default.err; tail -c 300 gpurun_out/r01b_bench_default.err
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r01b_bench_reference.json 2>> gpurun_out/r01b_bench_default.err
for wl in bunny_4k_full; do
  python bench.py --workload $wl --steps 150 --no-cpu-baseline > gpurun_out/r01b_bench_$wl.json 2>> gpurun_out/r01b_bench_default.err
done
bash tools/prof.sh r01b_smoke1080p_temporal smoke_1080p_temporal "k_primary|k_ris_thread"
bash tools/prof.sh r01b_bunny4k_full bunny_4k_full
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r01b_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("pass_ms"), "e2e", d.get("e2e", {}).get("value"), "roofline", d.get("roofline", {}).get("frac"), d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
