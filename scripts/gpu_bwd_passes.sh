#!/bin/bash
# Per-pass durations of the tensor-core backward (ncu launch list; cold-cache, serialised: read the SHARES).
set -u
TAG=${1:-bwd}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu -k "graph" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv \
    --log-file $OUT/launches_bwd.csv python scripts/bwd_tc_only.py > $OUT/launches_bwd.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("$OUT/launches_bwd.csv") if l.startswith('"')))
hdr = rows[0]; ik = hdr.index("Kernel Name"); im = hdr.index("Metric Name"); iv = hdr.index("Metric Value"); iid = hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[iid], r[ik][:40]), {})[r[im]] = float(r[iv].replace(",", ""))
for (i, k), m in per.items():
    print(i, k, " ".join(f"{a.split('__')[-1]}={v:.4g}" for a, v in m.items()))
PY
