#!/bin/bash
# One gpurun call: parity tests, short benches, block-kernel timings and one ncu capture per headline workload.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
timeout 600 python tools/run_once.py fateman1 pearce1 fateman2 monagan5 monagan3 monagan1 --iters 6 > gpurun_out/once.log 2>&1; cat gpurun_out/once.log
timeout 600 python tools/run_once.py fateman1 pearce1 fateman2 monagan5 --iters 6 --algo zone > gpurun_out/once_zone.log 2>&1; cat gpurun_out/once_zone.log
for w in pearce1 fateman1; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline; done > gpurun_out/bench3.log 2>&1
for w in fateman1 pearce1; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_block -s 1 -c 1 -f -o gpurun_out/block_$w python tools/run_once.py $w --iters 2 > gpurun_out/ncu_$w.log 2>&1
done
tail -2 gpurun_out/ncu_pearce1.log
