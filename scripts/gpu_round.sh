#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + one full capture of the top kernel.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh [tag]
set -u
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench cfg4" ; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -3 | tee $OUT/bench_cfg4.json
echo "== bench cfg3" ; timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_cfg3.json
echo "== bench cfg5" ; timeout 300 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_cfg5.json
echo "== bench cfg4 fp32 kernel" ; timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --precision fp32 2>&1 | tail -1 | tee $OUT/bench_cfg4_fp32.json
echo "== bench cfg2" ; timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_cfg2.json
echo "== bench cfg1" ; timeout 300 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_cfg1.json
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --batch 8192 --no-cpu > $OUT/launches_bench.log 2>&1
echo "== ncu full capture of the forward kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_forward -s 2 -c 1 -o $OUT/prof_fwd \
    python bench.py --steps 1 --warmup 1 --batch 8192 --no-cpu > $OUT/prof_bench.log 2>&1
echo "== ncu captures of the backward passes"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cc_dgrad_tc -s 1 -c 1 -o $OUT/prof_dgrad \
    python scripts/bwd_tc_bringup.py cfg3 > $OUT/prof_dgrad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cc_wgrad_tc -s 1 -c 1 -o $OUT/prof_wgrad \
    python scripts/bwd_tc_bringup.py cfg3 > $OUT/prof_wgrad.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $OUT/launches_bwd.csv \
    python scripts/bwd_tc_bringup.py cfg3 > $OUT/launches_bwd.log 2>&1
ls -la $OUT
