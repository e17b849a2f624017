#!/bin/bash
# Sweep of the backward's chunk size (tiles of 128 rows per CTA and chunk, UMNN_B200_BWD_TILES) on one box.
set -u
OUT=gpurun_out/${1:-tiles}
mkdir -p $OUT
for rep in 1 2; do
for t in 32 48 64 96 128; do
  export UMNN_B200_BWD_TILES=$t
  for sh in cfg3 cfg5 cfg2 cfg4m; do timeout 300 python scripts/bwd_time.py $sh 10 2>&1 | tail -1 | cut -c1-88 | sed "s/^/tiles=$t /"; done
done
done | tee $OUT/tiles.txt
unset UMNN_B200_BWD_TILES
echo "== flow bench (default)"; timeout 900 python scripts/flow_bench.py power mnist --no-torch 2>&1 | grep "^{" | cut -c1-330
