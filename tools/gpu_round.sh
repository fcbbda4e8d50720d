#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log; tail -n 4 gpurun_out/pytest.log
timeout 600 python tools/run_once.py fateman1 pearce1 fateman2 monagan5 monagan3 monagan1 --iters 6 2>&1 > gpurun_out/once.log; cut -c1-270 gpurun_out/once.log
for hz in 18 36; do echo "== block_hz=$hz"; timeout 600 python tools/run_once.py pearce1 --iters 5 block_hz=$hz 2>&1 | cut -c1-270; done
