#!/bin/bash
# diagnostics in one box: association tests + B=64 timings / ncu (PAF change), role counters of every conv launch in a step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_assoc_gpu.py tests/test_f4_gpu.py -x -q > gpurun_out/diag_pytest_assoc.log 2>&1; echo "pytest assoc rc=$?" > gpurun_out/diag_summary.txt
timeout 300 python tools/assoc_bw.py > gpurun_out/assoc.log 2>&1; echo "assoc rc=$?" >> gpurun_out/diag_summary.txt
timeout 300 ncu --profile-from-start off -k regex:'nms|paf|group' --clock-control none \
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
   --csv --log-file gpurun_out/assoc_ncu.csv python tools/assoc_bw.py --ncu > gpurun_out/assoc_ncu.log 2>&1; echo "assoc ncu rc=$?" >> gpurun_out/diag_summary.txt
SMAPB_ROLES_PLAN=gpurun_out/roles_plan.csv timeout 300 python tools/roles_plan.py > gpurun_out/roles_plan.txt 2> gpurun_out/roles_plan.err; echo "roles rc=$?" >> gpurun_out/diag_summary.txt
cat gpurun_out/diag_summary.txt; tail -3 gpurun_out/diag_pytest_assoc.log; cat gpurun_out/assoc.log; grep -a "paf\|nms\|group" gpurun_out/assoc_ncu.csv | cut -d, -f5,11- | head -20; head -60 gpurun_out/roles_plan.txt; tail -5 gpurun_out/roles_plan.err
