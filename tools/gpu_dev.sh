#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PK_ONCE_ALL=1 timeout 120 python tools/run_once.py pearce1 fateman1 --iters 8 2>&1 | tail -4
PK_ONCE_ALL=1 timeout 120 python tools/run_once.py pearce1 --iters 8 block_ctas=2 2>&1 | tail -2
timeout 60 python tools/run_once.py pearce1 --iters 6 --part 0/2 2>&1 | tail -1
timeout 60 python tools/run_once.py pearce1 --iters 6 --part 1/2 2>&1 | tail -1
timeout 60 python tools/run_once.py pearce1 --iters 6 --part 3/8 2>&1 | tail -1
timeout 60 python tools/run_once.py pearce1 --iters 6 --part 3/8 block_ctas=2 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_once.csv python tools/run_once.py pearce1 --iters 3 > gpurun_out/ncu_l.log 2>&1
