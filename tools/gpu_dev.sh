#!/bin/bash
cd "$(dirname "$0")/.."
echo "--- part 0/2"
PK_ONCE_ALL=1 timeout 120 python tools/run_once.py pearce1 --iters 12 --part 0/2 2>&1 | tail -2 | head -1
echo "--- part 3/8"
PK_ONCE_ALL=1 timeout 120 python tools/run_once.py pearce1 --iters 12 --part 3/8 2>&1 | tail -2 | head -1
echo "--- whole"
PK_ONCE_ALL=1 timeout 120 python tools/run_once.py pearce1 fateman1 --iters 10 2>&1 | grep device_ms:
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --no-cpu-baseline --also fateman1 2>/dev/null | python -c "
import json,sys
d=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('bench ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['also']['fateman1']['ms_per_step'], d['step_ms_rank0'])"
