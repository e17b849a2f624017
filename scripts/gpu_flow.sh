#!/bin/bash
# Flow-level GPU visit: new parity tests, then compute_ll / training-step / invert timings per driver shape.
set -u
TAG=${1:-flow}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest (invert, graphs, flow)"; timeout 600 python -m pytest tests -x -q -m gpu -k "invert or graph or flow" 2>&1 | tail -15 | tee $OUT/pytest_flow.txt
echo "== flow bench"; timeout 1500 python scripts/flow_bench.py toy power mnist bsds 2>&1 | grep -v Warning | tail -12 | tee $OUT/flow_bench.jsonl
