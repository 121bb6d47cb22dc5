#!/bin/bash
# ncu launch list (durations only) of the kernels matching $2 over a 2-step bench run
mkdir -p gpurun_out
TAG=${1:-l}
KPAT=${2:-k_tile}
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k "regex:${KPAT}" -c 60 --csv --log-file gpurun_out/${TAG}_list.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_list.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_list.csv")) if len(r)>10 and r[0].isdigit()]
from collections import OrderedDict
d=OrderedDict()
for r in rows:
    d.setdefault(r[0],{"k":r[4][:70]})[r[12]]=r[14]
for k,v in d.items():
    print(k, v["k"], " ".join(f"{a.split('__')[-1][:18]}={b}" for a,b in v.items() if a!="k"))
PY
