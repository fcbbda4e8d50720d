#!/bin/bash
# scratch: launch list of one product (edit freely)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
W=${1:-fateman1}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/dev_launches.csv python tools/run_once.py $W --iters 3 > gpurun_out/dev_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/dev_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
names=[(r[ki][:60], float(r[vi].replace(",",""))) for r in rows[1:]]
last=max(i for i,(n,_) in enumerate(names) if "k_prepare" in n)
for n,v in names[last:]: print(f"{v/1000:9.2f} us  {n}")
print("sum", sum(v for _,v in names[last:])/1000)
PY
