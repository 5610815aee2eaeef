#!/bin/bash
# short end-of-session check at HEAD: all -m gpu tests, smoke, bench (both arms)
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
timeout 1500 python -m pytest tests -q -m gpu -x -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $S
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 --profile-csv gpurun_out/ops.csv > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> $S
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> $S
cat $S; grep -a "e2e \|passed\|failed" gpurun_out/pytest_gpu.log | tail -4; tail -1 gpurun_out/smoke.log; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f ms %.3f clocks %s frac %.3f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"], d["roofline"]["frac"], d["gpu_launches"]))
PY
