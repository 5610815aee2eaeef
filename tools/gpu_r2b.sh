#!/bin/bash
# round 2, second pass: all -m gpu tests (new: e2e chain, sharding invariance, bit-exact tiles), association at B=64 (events + ncu)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary.txt
timeout 300 python tools/assoc_bw.py > gpurun_out/assoc.log 2>&1; echo "assoc rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --profile-from-start off -k regex:'nms|paf|group' --clock-control none \
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
   --csv --log-file gpurun_out/assoc_ncu.csv python tools/assoc_bw.py --ncu > gpurun_out/assoc_ncu.log 2>&1; echo "assoc ncu rc=$?" >> gpurun_out/summary.txt
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -a "e2e \|passed\|failed\|Error" gpurun_out/pytest_gpu.log | tail -12; tail -30 gpurun_out/pytest_gpu.log | head -60; cat gpurun_out/assoc.log; grep -a "nms\|paf\|group" gpurun_out/assoc_ncu.csv | cut -c1-300 | tail -16; tail -3 gpurun_out/smoke.log
