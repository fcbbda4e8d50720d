#!/bin/bash
cd "$(dirname "$0")/.."
timeout 120 python tools/run_once.py fateman1 fateman2 monagan5 --iters 8 2>&1 | tail -3 | cut -c1-230
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/r2_launches_warm_f1.csv python tools/run_once.py fateman1 --iters 6 > gpurun_out/ncu_l.log 2>&1
grep -c gather gpurun_out/r2_launches_warm_f1.csv; grep gather gpurun_out/r2_launches_warm_f1.csv | tail -1 | cut -c1-60; grep gather gpurun_out/r2_launches_warm_f1.csv | tail -1 | rev | cut -d, -f1 | rev
