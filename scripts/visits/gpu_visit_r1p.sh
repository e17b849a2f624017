#!/bin/bash
# GPU visit: CUDA-graph replay of the invert refinement loop -- parity suite and flow-level invert timings.
set -u
TAG=${1:-r1p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== flow bench (graph)"; timeout 900 python scripts/flow_bench.py --no-torch 2>&1 | grep "^{" | tee $OUT/flow_bench_graph.jsonl | cut -c1-330
echo "== flow bench (no graph)"; UMNN_B200_INVERT_GRAPH=0 timeout 900 python scripts/flow_bench.py --no-torch 2>&1 | grep "^{" | tee $OUT/flow_bench_nograph.jsonl | cut -c1-330
echo "== bench cfg4"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_cfg4.json | cut -c1-200
