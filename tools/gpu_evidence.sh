#!/bin/bash
# One gpurun call that refreshes the judged evidence for a kernel revision:
#   gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh r1k'
# Order matters: every ncu pass runs only after the same command has exited 0 without ncu.
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
set -x
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -3 $O/pytest_gpu_$TAG.log
timeout 600 python bench.py > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_ref.json 2> $O/bench_${TAG}_ref.err; echo "ref rc=$?"
B3="python bench.py --steps 2 --warmup 1 --rays 1e9 --no-e2e --no-cpu-baseline"
B5="python bench.py --workload cfg5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-smoothing"
timeout 300 $B3 > $O/plain_${TAG}.log 2>&1 && {
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_${TAG}.csv $B3 > $O/ncu_l_${TAG}.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_exchange -s 2 -c 1 -f -o $O/prof_${TAG}_sq $B3 > $O/ncu_sq_${TAG}.log 2>&1
}
timeout 300 $B5 > $O/plain5_${TAG}.log 2>&1 && {
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_exchange -s 2 -c 1 -f -o $O/prof_${TAG}_queue $B5 > $O/ncu_q_${TAG}.log 2>&1
}
for w in cfg1 cfg2 cfg4 cfg5; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-smoothing > $O/bench_${TAG}_$w.json 2> $O/bench_${TAG}_$w.err
done
ls -la $O | tail -20
