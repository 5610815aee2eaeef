#!/bin/bash
mkdir -p gpurun_out
export SMAPB_NO_TILE_TABLE=1 SMAPB_NO_AUTOTUNE=1
for part in conv path; do
  timeout 900 compute-sanitizer --tool initcheck --print-limit 20 python tools/sanitize_targets.py $part > gpurun_out/sanitizer_initcheck_${part}.log 2>&1
  echo "initcheck(after zero-filling TMA-stored buffers) $part rc=$? : $(grep -a 'ERROR SUMMARY' gpurun_out/sanitizer_initcheck_${part}.log | tail -1)" >> gpurun_out/sanitizer_summary.txt
done
tail -3 gpurun_out/sanitizer_summary.txt
