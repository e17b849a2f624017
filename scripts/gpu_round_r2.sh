#!/bin/bash
# Round 2, full visit of the final state: parity suite, smoke, bench lines of every configuration and the reference arm, backward
# and flow timings, gradient flip analysis, launch lists, ncu full captures (wide / narrow forward, the three backward passes),
# compute-sanitizer memcheck.  Usage (under gpurun): bash scripts/gpu_round_r2.sh; then scripts/collect_r2_profiles.sh here.
set -u
OUT=gpurun_out/${1:-r2final}
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench cfg4 (default line)"; timeout 1200 python bench.py 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_cfg4.json | cut -c1-600
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_reference.json | cut -c1-400
for wl in cfg1 cfg2 cfg3 cfg5; do
  echo "== bench $wl"; timeout 400 python bench.py --workload $wl --steps 200 --warmup 20 --no-train 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_$wl.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], 'ms/step', round(d['ms_per_step'],4), 'G evals/s', round(d['value']/1e9,3), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'issued', round(d['roofline']['issued_tensor_frac'],3), 'err', d['parity']['integral_max_rel_err_vs_oracle'], 'x ref-cuda', round(d.get('aux',{}).get('speedup_vs_reference_torch_cuda',0),1))"
done
echo "== bench cfg4 fp32 kernel"; timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --no-train --precision fp32 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_cfg4_fp32.json | cut -c1-300
echo "== backward timing"
for sh in cfg3 cfg2 cfg5 cfg4s cfg4m; do timeout 300 python scripts/bwd_time.py $sh 2>&1 | tail -1 | tee -a $OUT/bwd_time.txt; done
UMNN_B200_BWD_PANELS=hilo timeout 300 python scripts/bwd_time.py cfg3 2>&1 | tail -1 | tee -a $OUT/bwd_time.txt
echo "== flow bench"; timeout 1200 python scripts/flow_bench.py 2>&1 | grep "^{" | tee $OUT/flow_bench.jsonl | cut -c1-330
echo "== grad flip analysis"; timeout 900 python scripts/grad_flip_analysis.py 2>&1 | tee $OUT/grad_flip_analysis.txt | grep -c "d_h"
echo "== ncu launch list (cfg4, B=8192)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --batch 8192 --no-cpu --no-train > $OUT/launches_bench.log 2>&1
echo "== ncu launch list backward (cfg3)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_bwd.csv \
    python scripts/bwd_tc_only.py > $OUT/launches_bwd.log 2>&1
echo "== ncu full capture of the forward kernel (cfg4, B=8192)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_forward_tc -s 2 -c 1 -o $OUT/prof_fwd \
    python bench.py --steps 1 --warmup 1 --batch 8192 --no-cpu --no-train > $OUT/prof_bench.log 2>&1
echo "== ncu full capture of the narrow forward kernel (cfg5)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_forward_tc -s 2 -c 1 -o $OUT/prof_fwd_cfg5 \
    python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu --no-train > $OUT/prof_bench_cfg5.log 2>&1
echo "== ncu captures of the backward passes (cfg3)"
for k in cc_forward_tc cc_dgrad_tc cc_wgrad_tc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $OUT/prof_bwd_$k \
      python scripts/bwd_tc_only.py > $OUT/prof_bwd_$k.log 2>&1
done
echo "== compute-sanitizer memcheck"
for which in fp32 fp16x3 fp16x3_hi overflow invert; do
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_case.py $which 2>&1 | grep -vE "^$" | tail -6 | tee $OUT/memcheck_$which.txt
done
ls -la $OUT
