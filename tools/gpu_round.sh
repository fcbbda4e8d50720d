#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for pp in 0 2 4 8; do echo "== pipe_parts=$pp"; PK_PIPE_PARTS=$pp timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"e2e": {[^}]*}'; done
build/bin/pearce1 3 2>&1 | tail -n 2
timeout 900 python -m pytest tests -m gpu -x -q -k "pipelined" 2>&1 | tail -n 3
