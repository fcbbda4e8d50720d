#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PK_ONCE_ALL=1 timeout 120 python tools/run_once.py pearce1 fateman1 pearce2 --iters 8 2>&1 | tail -6 | cut -c1-230
timeout 60 python tools/run_once.py pearce1 --iters 6 block_size_hint=0 2>&1 | tail -1 | cut -c1-230
timeout 60 python tools/run_once.py pearce1 --iters 6 --part 3/8 2>&1 | tail -1 | cut -c1-230
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
