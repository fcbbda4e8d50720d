#!/bin/bash
# Scaling bench on the GPUs of one box (run under `gpurun --gpus N`): pearce1 (headline) at the rank counts given ($@, default 1 2 4 8).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
W=${PK_SCALE_WORKLOAD:-pearce1}
for N in ${@:-1 2 4 8}; do
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --workload $W --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/r2_scale_${W}_1.log 2> gpurun_out/r2_scale_${W}_1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --workload $W --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_scale_${W}_$N.log 2> gpurun_out/r2_scale_${W}_$N.err
  fi
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2_scale_${W}_$N.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("N=$N ms/step %.3f value %.3g e2e ms %.3f parity %s | per-rank %s"%(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d.get("parity_ok"), [ (round(r["device_ms"],3), round(r["kernel_ms"],3)) for r in d.get("details",{}).get("per_rank",[])]))
    if "e2e_per_rank_parts" in d: print("   sharded e2e ms %.3f"%d["e2e_per_rank_parts"]["ms_per_step"], d["e2e"].get("api"))
except Exception as e:
    print("N=$N failed", e); print(open("gpurun_out/r2_scale_${W}_$N.err").read()[-1500:])
PY
done
