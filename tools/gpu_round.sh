#!/bin/bash
# Final artefacts of a round: bench (both arms), launch list, one full ncu capture per headline workload, microbenchmarks, C++ programs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${PK_ROUND:-r2}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python bench.py > gpurun_out/${R}_bench_pearce1.json 2> gpurun_out/${R}_bench_pearce1.err; tail -c 400 gpurun_out/${R}_bench_pearce1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${R}_bench_reference_pearce1.json 2>&1; tail -c 300 gpurun_out/${R}_bench_reference_pearce1.json
timeout 300 python tools/run_once.py fateman1 pearce1 fateman2 monagan1 monagan2 monagan3 monagan4 monagan5 pearce2 fateman1_dynamic pearce1_dynamic --iters 8 2>&1 | grep -v "^block path" > gpurun_out/${R}_workloads.txt
timeout 100 python tools/run_once.py fateman1 --iters 6 --trunc 20 >> gpurun_out/${R}_workloads.txt 2>&1; timeout 100 python tools/run_once.py fateman1 --iters 6 --trunc 30 >> gpurun_out/${R}_workloads.txt 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/plain_l.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_pearce1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/ncu_l.log 2>&1
timeout 200 python tools/run_once.py pearce1 --iters 2 > gpurun_out/plain_p.log 2>&1 && timeout 300 ncu --set full --clock-control none -k regex:k_block_acc -s 1 -c 1 -f -o gpurun_out/${R}_acc_pearce1 python tools/run_once.py pearce1 --iters 2 > gpurun_out/ncu_acc.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_block_mark -s 1 -c 1 -f -o gpurun_out/${R}_mark_pearce1 python tools/run_once.py pearce1 --iters 2 > gpurun_out/ncu_mark.log 2>&1
python tools/ncu_summary.py gpurun_out/${R}_mark_pearce1.ncu-rep > gpurun_out/${R}_pearce1_k_block_mark.txt 2>&1; rm -f gpurun_out/${R}_mark_pearce1.ncu-rep  # (three reports exceed what gpurun copies back)
timeout 200 python tools/run_once.py fateman1 --iters 2 > gpurun_out/plain_f.log 2>&1 && timeout 300 ncu --set full --clock-control none -k regex:k_block_dense -s 1 -c 1 -f -o gpurun_out/${R}_dense_fateman1 python tools/run_once.py fateman1 --iters 2 > gpurun_out/ncu_dense.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench tools/microbench.cu > /dev/null 2>&1; timeout 300 tools/microbench > gpurun_out/${R}_microbench.txt 2>&1
(timeout 100 build/bin/fateman1 5 | tail -n 1; timeout 100 build/bin/fateman2 3 | tail -n 1; timeout 100 build/bin/pearce1 5 | tail -n 1; timeout 100 build/bin/fateman1_rational 5 | tail -n 1; timeout 100 build/bin/pearce1_rational 3 | tail -n 1) > gpurun_out/${R}_cpp_benchmarks.txt 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
