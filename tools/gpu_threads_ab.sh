#!/bin/bash
# A/B of the SQ block shape (RTHX_SQ_THREADS) on cfg3 at 1e9 rays: 256 (64 regs, 32 warps/SM), 288 (56, 36), 320 (48, 40)
for T in 256 288 320 256 288 320; do
  RTHX_SQ_THREADS=$T python bench.py --steps 10 --warmup 3 --rays 1e9 --no-e2e --no-cpu-baseline --no-smoothing > gpurun_out/ab_$T.json 2>gpurun_out/ab_$T.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$T.json").read().strip().splitlines()[-1])
print($T, "%.4g rays/s" % d["value"], "%.3f ms" % d["ms_per_step"], d["check"], d["config"]["launch"])
PY
done
