#!/bin/bash
# Multi-GPU bench on N GPUs of one box (run under `gpurun --gpus N`): N=1 then N=$1, pearce1 and fateman1.
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
for w in pearce1 fateman2; do
  timeout 600 python bench.py --workload $w --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${w}_1.log 2>&1; tail -c 900 gpurun_out/scale_${w}_1.log | cut -c1-900
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --workload $w --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_${w}_$N.log 2>&1; tail -c 1500 gpurun_out/scale_${w}_$N.log | cut -c1-1500
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/scale_ref_$N.log 2>&1; tail -c 600 gpurun_out/scale_ref_$N.log
