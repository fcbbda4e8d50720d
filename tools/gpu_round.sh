#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*\|"device_ms": [0-9.]*\|"gpu_launches": [0-9]*'
PK_FINAL_SYNC=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*, "higher\|"kernel_ms": [0-9.]*\|"device_ms": [0-9.]*\|"gpu_launches": [0-9]*'
