#!/bin/bash
# end-of-round GPU pass in ONE box: all -m gpu tests, smoke, bench (both arms), per-op CSV, ncu launch list of the bench step,
# association at B = 64 (events + ncu)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary.txt
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 --profile-csv gpurun_out/ops.csv > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> gpurun_out/summary.txt
SMAPB_NO_GRAPH=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --ncu-one-step --warmup 3 --engines 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu launch list rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/assoc_bw.py > gpurun_out/assoc.log 2>&1; echo "assoc rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --profile-from-start off -k regex:'nms|paf|group' --clock-control none \
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
   --csv --log-file gpurun_out/assoc_ncu.csv python tools/assoc_bw.py --ncu > gpurun_out/assoc_ncu.log 2>&1; echo "assoc ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -a "e2e \|passed\|failed" gpurun_out/pytest_gpu.log | tail -5; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_ref.json; cat gpurun_out/assoc.log
