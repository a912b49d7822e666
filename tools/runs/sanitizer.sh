#!/bin/bash
# compute-sanitizer over the round-2 kernels (small inputs); output -> gpurun_out/r02_compute_sanitizer.txt
mkdir -p gpurun_out; O=gpurun_out/r02_compute_sanitizer.txt; : > $O
for tool in memcheck racecheck; do for w in fit solve xchg b16; do
  echo "== $tool $w" >> $O
  timeout 900 compute-sanitizer --tool $tool python tests/tools/sanitize_small.py $w 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|done|xchg|fit|tranche|b16" | head -20 >> $O
done; done
cat $O
