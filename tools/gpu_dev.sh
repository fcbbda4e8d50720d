#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/run_once.py fateman1 pearce1 monagan1 monagan5 --iters 8 2>&1 | grep -v "^block path" | cut -c1-230
timeout 600 python -m pytest tests -x -q -m gpu -k "fateman or pearce or monagan or golden or random or limb" 2>&1 | tail -4
