#!/bin/bash
# compute-sanitizer over the small-shape workload; logs -> gpurun_out/sanitizer_*.log
# (initcheck cannot see TMA stores: the library zero-fills TMA-stored buffers once at allocation so that it stays clean)
mkdir -p gpurun_out
export SMAPB_NO_TILE_TABLE=1 SMAPB_NO_AUTOTUNE=1
for tool in memcheck synccheck initcheck racecheck; do
  for part in conv path assoc; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_targets.py $part > gpurun_out/sanitizer_${tool}_${part}.log 2>&1
    echo "$tool $part rc=$? : $(grep -a 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_${part}.log | tail -1)" >> gpurun_out/sanitizer_summary.txt
  done
done
cat gpurun_out/sanitizer_summary.txt
