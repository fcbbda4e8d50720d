#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log; tail -n 25 gpurun_out/pytest.log | cut -c1-200
timeout 600 python tools/run_once.py fateman1 pearce1 fateman2 monagan5 --iters 6 2>&1 > gpurun_out/once.log; cut -c1-270 gpurun_out/once.log
for D in 20 30; do echo "== truncated D=$D block / dense"; timeout 600 python tools/run_once.py fateman1 --iters 5 --trunc $D 2>&1 | cut -c1-270; timeout 600 python tools/run_once.py fateman1 --iters 5 --trunc $D --algo dense 2>&1 | cut -c1-270; done
