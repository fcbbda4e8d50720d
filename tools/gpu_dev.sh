#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/run_once.py pearce1 pearce2 pearce1_dynamic --iters 6 2>&1 | grep -v "^block path" | cut -c1-230
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
