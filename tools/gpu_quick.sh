#!/bin/bash
# Quick check of a kernel change: GPU parity tests + short benches (cfg3 at 1e9 rays, cfg5).  gpurun --timeout 900 -- 'bash tools/gpu_quick.sh tag'
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu_$TAG.log
for w in cfg3 cfg5 cfg2; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --rays 1e9 --no-e2e --no-cpu-baseline --no-smoothing > $O/bench_${TAG}_$w.json 2> $O/bench_${TAG}_$w.err
  python - <<PY
import json
d=json.loads(open("$O/bench_${TAG}_$w.json").read().strip().splitlines()[-1])
print("$w", "%.4g rays/s" % d["value"], "%.3f ms" % d["ms_per_step"], d["check"], d["config"]["launch"])
PY
done
