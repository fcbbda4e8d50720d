#!/bin/bash
# Final artefacts of a round: bench (both arms), launch list, one full ncu capture per headline workload.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_pearce1.log 2>&1; tail -c 3000 gpurun_out/bench_pearce1.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; tail -c 900 gpurun_out/bench_ref.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_l.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_pearce1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
for w in fateman1 pearce1; do
  timeout 600 python tools/run_once.py $w --iters 2 > gpurun_out/plain_$w.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k k_block -s 1 -c 1 -f -o gpurun_out/block_$w python tools/run_once.py $w --iters 2 > gpurun_out/ncu_$w.log 2>&1
done
tail -n 2 gpurun_out/ncu_pearce1.log
