#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log; tail -n 4 gpurun_out/pytest.log
timeout 600 python tools/run_once.py fateman1 pearce1 fateman2 monagan5 monagan3 monagan1 --iters 6 2>&1 > gpurun_out/once.log; cut -c1-270 gpurun_out/once.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pearce1.log 2>&1; tail -c 1800 gpurun_out/bench_pearce1.log
for w in fateman1 pearce1; do
  timeout 600 ncu --set full --clock-control none --import-source on -k k_block -s 1 -c 1 -f -o gpurun_out/block_$w python tools/run_once.py $w --iters 2 > gpurun_out/ncu_$w.log 2>&1
done
tail -n 2 gpurun_out/ncu_pearce1.log
