#!/bin/bash
# Round 2, visit k: ncu --set full captures after the prep-warp rebuild (wide forward cfg4/8192, narrow forward cfg5, passes F and D).
set -u
OUT=gpurun_out/${1:-r2k}
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_forward_tc -s 2 -c 1 -o $OUT/prof_fwd \
    python bench.py --steps 1 --warmup 1 --batch 8192 --no-cpu --no-train > $OUT/prof_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_forward_tc -s 2 -c 1 -o $OUT/prof_fwd_cfg5 \
    python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu --no-train > $OUT/prof_bench_cfg5.log 2>&1
for k in cc_forward_tc cc_dgrad_tc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $OUT/prof_bwd_$k \
      python scripts/bwd_tc_only.py > $OUT/prof_bwd_$k.log 2>&1
done
ls -la $OUT
