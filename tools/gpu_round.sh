#!/bin/bash
# full GPU check: all -m gpu tests, smoke, short bench (+ per-op CSV)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary.txt
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --profile-csv gpurun_out/ops.csv > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -15 gpurun_out/pytest_gpu.log; tail -5 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
