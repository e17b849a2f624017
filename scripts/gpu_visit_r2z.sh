#!/bin/bash
# Round 2, visit z: compute-sanitizer (memcheck + racecheck) on a multi-tile case: the producer / consumer named barriers of the
# finalize step and of pass D's d_h reduction, the prep-buffer ring and the carried partial sums only act from the second tile on.
set -u
OUT=gpurun_out/${1:-r2z}
mkdir -p $OUT
for tool in memcheck racecheck; do
  echo "=== $tool multitile"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_case.py multitile > $OUT/${tool}_multitile_full.txt 2>&1
  echo "rc=$?"
  grep -E "Error: Race|and (Read|Write) access|SUMMARY|multitile:|hazards\]" $OUT/${tool}_multitile_full.txt | sed -E 's/_ZN4umnn[0-9A-Za-z_]*cc_/cc_/' | cut -c1-230 | sort | uniq -c | sort -rn | head -40 | tee $OUT/${tool}_multitile.txt
done
