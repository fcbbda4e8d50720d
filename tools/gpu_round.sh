#!/bin/bash
# Final artefacts of a round: bench (both arms), launch list, one full ncu capture per headline workload, C++ programs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 600 python bench.py > gpurun_out/bench_pearce1.log 2>&1; tail -c 3600 gpurun_out/bench_pearce1.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -c 700 gpurun_out/bench_ref.log
timeout 300 python bench.py --workload fateman1 --no-cpu-baseline > gpurun_out/bench_fateman1.log 2>&1
timeout 300 python bench.py --workload fateman2 --no-cpu-baseline > gpurun_out/bench_fateman2.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_l.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_pearce1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
for w in fateman1 pearce1; do
  timeout 200 python tools/run_once.py $w --iters 2 > gpurun_out/plain_$w.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k k_block -s 1 -c 1 -f -o gpurun_out/block_$w python tools/run_once.py $w --iters 2 > gpurun_out/ncu_$w.log 2>&1
done
timeout 200 python tools/run_once.py fateman1 pearce1 fateman2 monagan1 monagan2 monagan3 monagan4 monagan5 pearce2 fateman1_dynamic pearce1_dynamic --iters 8 > gpurun_out/once_all.log 2>&1
timeout 100 python tools/run_once.py fateman1 --iters 6 --trunc 20 >> gpurun_out/once_all.log 2>&1; timeout 100 python tools/run_once.py fateman1 --iters 6 --trunc 30 >> gpurun_out/once_all.log 2>&1
(timeout 100 build/bin/fateman1 5 | tail -n 1; timeout 100 build/bin/fateman2 3 | tail -n 1; timeout 100 build/bin/pearce1 5 | tail -n 1; timeout 100 build/bin/fateman1_rational 5 | tail -n 1; timeout 100 build/bin/pearce1_rational 3 | tail -n 1) > gpurun_out/cpp_bench.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
