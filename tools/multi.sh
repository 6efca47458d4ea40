#!/bin/bash
# usage: tools/multi.sh <N> <workload> <steps> [tag]  -> gpurun_out/r01b_bench_<tag>_n<N>.json
N=$1; wl=$2; steps=$3; tag=${4:-$wl}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --workload $wl --steps $steps --warmup 8 \
  > gpurun_out/r01b_bench_${tag}_n$N.json 2> gpurun_out/multi_err_${tag}_n$N.log
python - "$N" "$tag" <<'PY'
import json, sys
f = "gpurun_out/r01b_bench_%s_n%s.json" % (sys.argv[2], sys.argv[1])
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(d["n_gpus"], d["config"]["name"], d["value"], d["ms_per_step"], d["pass_ms"], "e2e", d["e2e"]["value"], d["config"]["partition"], d.get("per_rank_initial_ms"))
except Exception as e:
    print(f, "failed", e); print(open("gpurun_out/multi_err_%s_n%s.log" % (sys.argv[2], sys.argv[1])).read()[-1500:])
PY
