#!/bin/bash
# Turn the files a `scripts/gpu_round_r2.sh` visit merged into gpurun_out/r2final into the tracked evidence under profiles/.
set -u
SRC=gpurun_out/${1:-r2final}
DST=profiles
for wl in cfg1 cfg2 cfg3 cfg4 cfg5; do cp $SRC/bench_$wl.json $DST/r2_bench_$wl.json; done
cp $SRC/bench_cfg4_fp32.json $DST/r2_bench_cfg4_fp32.json
cp $SRC/bench_reference.json $DST/r2_bench_reference.json
cp $SRC/bwd_time.txt $DST/r2_bwd_time.txt
cp $SRC/flow_bench.jsonl $DST/r2_flow_bench.jsonl
cp $SRC/grad_flip_analysis.txt $DST/r2_grad_flip_analysis.txt
cp $SRC/pytest_gpu.txt $DST/r2_pytest_gpu.txt
cp $SRC/launches.csv $DST/r2_launches.csv
cp $SRC/launches_bwd.csv $DST/r2_launches_bwd.csv
mkdir -p $DST/sanitizer
for f in $SRC/memcheck_*.txt; do cp $f $DST/sanitizer/r2_$(basename $f); done
python scripts/ncu_summary.py $SRC/prof_fwd.ncu-rep $DST/r2_fwd_wide_cfg4_ncu_summary.csv
python scripts/ncu_summary.py $SRC/prof_fwd_cfg5.ncu-rep $DST/r2_fwd_narrow_cfg5_ncu_summary.csv
python scripts/ncu_summary.py $SRC/prof_bwd_cc_forward_tc.ncu-rep $DST/r2_bwd_passF_ncu_summary.csv
python scripts/ncu_summary.py $SRC/prof_bwd_cc_dgrad_tc.ncu-rep $DST/r2_bwd_passD_ncu_summary.csv
python scripts/ncu_summary.py $SRC/prof_bwd_cc_wgrad_tc.ncu-rep $DST/r2_bwd_passW_ncu_summary.csv
python scripts/sass_summary.py > $DST/sass_summary.txt
python - <<'PY'
import csv, json
def dram(path):
    rd = wr = None
    for m, u, v in csv.reader(open(path)):
        if m == "dram__bytes_read.sum": rd = float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
        if m == "dram__bytes_write.sum": wr = float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
    return rd + wr
t = json.load(open("profiles/traffic.json"))
b = dram("profiles/r2_fwd_wide_cfg4_ncu_summary.csv")
t["cfg4:tc"] = {"dram_bytes_per_slot": b / (8192 * 63),
                "source": "profiles/r2_fwd_wide_cfg4_ncu_summary.csv (ncu --set full, B=8192 of cfg4: dram__bytes_read.sum + dram__bytes_write.sum)"}
b5 = dram("profiles/r2_fwd_narrow_cfg5_ncu_summary.csv")
t["cfg5:tc"] = {"dram_bytes_per_slot": b5 / (100 * 784),
                "source": "profiles/r2_fwd_narrow_cfg5_ncu_summary.csv (ncu --set full, cfg5)"}
json.dump(t, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(t, indent=1))
PY
