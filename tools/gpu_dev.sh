#!/bin/bash
# scratch: A/B of a tuning key on the GPU box (edit freely)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ilp in 0 1 2 3; do
  echo "== block_ilp=$ilp"
  timeout 200 python tools/run_once.py pearce1 pearce2 monagan3 pearce1_dynamic --iters 8 block_ilp=$ilp 2>&1 | grep -v "^block path" | cut -c1-260
done
for ilp in 0 3; do
  echo "== block_ilp=$ilp (dense family, must not change)"
  timeout 200 python tools/run_once.py fateman1 --iters 6 block_ilp=$ilp 2>&1 | cut -c1-260
done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
