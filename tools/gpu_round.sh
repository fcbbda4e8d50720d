#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PK_BLOCK_DEBUG=1
for a in "" "block_radix0=32" "block_hz=8" "block_hz=24" "block_radix0=32 block_hz=8" "block_mode=1"; do echo "== $a"; timeout 600 python tools/run_once.py pearce1 --iters 3 block_target=1024 $a 2>&1 | sort -u; done > gpurun_out/once_p.log 2>&1; cut -c1-260 gpurun_out/once_p.log
unset PK_BLOCK_DEBUG
timeout 600 ncu --set full --clock-control none --import-source on -k k_block -s 1 -c 1 -f -o gpurun_out/block_pearce1_r0 python tools/run_once.py pearce1 --iters 2 block_target=1024 > gpurun_out/ncu_p0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k k_block -s 1 -c 1 -f -o gpurun_out/block_pearce1_r32 python tools/run_once.py pearce1 --iters 2 block_target=1024 block_radix0=32 > gpurun_out/ncu_p32.log 2>&1
tail -2 gpurun_out/ncu_p32.log
