#!/bin/bash
# Scaling bench on the GPUs of one box (run under `gpurun --gpus N`): pearce1 (headline) and fateman2 at N = 1, 2, 4, ... up to $1.
cd "$(dirname "$0")/.."
NMAX=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
for w in pearce1 fateman2; do
  for N in 1 2 4 8; do
    [ $N -gt $NMAX ] && continue
    if [ $N -eq 1 ]; then
      timeout 600 python bench.py --workload $w --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${w}_1.log 2>&1
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --workload $w --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_${w}_$N.log 2>&1
    fi
    grep -o '"n_gpus": [0-9]*\|"ms_per_step": [0-9.]*, "higher\|"value": [0-9.e+]*, "unit": "term-products/s", "n_gpus"\|"kernel_ms": [0-9.]*\|"device_ms": [0-9.]*' gpurun_out/scale_${w}_$N.log | tr '\n' ' '; echo " <- $w N=$N"
  done
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29599 bench.py --impl reference --gpus $NMAX --steps 2 --warmup 1 > gpurun_out/scale_ref_$NMAX.log 2>&1; tail -c 400 gpurun_out/scale_ref_$NMAX.log
