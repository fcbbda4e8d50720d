#!/bin/bash
# scratch: timing of the headline workloads (edit freely)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/run_once.py pearce1 pearce2 pearce1_dynamic --iters 8 2>&1 | grep -v "^block path" | cut -c1-260
timeout 100 python tools/run_once.py pearce1 --iters 6 --part 3/8 2>&1 | cut -c1-260
timeout 100 python tools/run_once.py pearce1 --iters 6 --part 0/2 2>&1 | cut -c1-260
timeout 600 python -m pytest tests -x -q -m gpu -k "block or pearce or trunc or multi_ctx" 2>&1 | tail -2
