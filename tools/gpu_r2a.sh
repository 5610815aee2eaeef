#!/bin/bash
# round 2, first pass: new CTA-pair tile shapes (bits + speed), tile table, tests, bench with / without serpentine order
mkdir -p gpurun_out
timeout 300 python tools/tile_probe.py > gpurun_out/tile_probe.txt 2>&1; echo "probe rc=$?" > gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu -x > gpurun_out/pytest_conv.log 2>&1; echo "pytest conv rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/make_tile_table.py gpurun_out/b200.tsv > gpurun_out/tile_table.log 2>&1; echo "table rc=$?" >> gpurun_out/summary.txt
mkdir -p smap_b200/tiles; cp gpurun_out/b200.tsv smap_b200/tiles/b200.tsv
timeout 900 python -m pytest tests/test_backbone_gpu.py tests/test_pipeline_gpu.py tests/test_shims_gpu.py -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-csv gpurun_out/ops.csv > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
SMAPB_SERPENTINE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-csv gpurun_out/ops_serp.csv > gpurun_out/bench_serp.json 2> gpurun_out/bench_serp.err; echo "bench serp rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --engines 3 > gpurun_out/bench_e3.json 2> gpurun_out/bench_e3.err; echo "bench e3 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/tile_probe.txt; tail -5 gpurun_out/pytest_conv.log; tail -5 gpurun_out/tile_table.log; tail -8 gpurun_out/pytest_gpu.log
for f in bench bench_serp bench_e3; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$f.json")); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], d["roofline"]["frac"], d["breakdown_ms_per_step"])
except Exception as e: print("$f failed", e); print(open("gpurun_out/$f.err").read()[-1500:])
PY
done
