#!/bin/bash
# A/B of a kernel change on ONE box: conv parity tests with the new build, output digests of both builds (must be identical
# when the change claims to keep the arithmetic), bench and single-layer timings of both builds back to back.
# The baseline build is smap_b200/lib/libsmap_b200_base.so (a copy of the previous libsmap_b200.so).
mkdir -p gpurun_out
BASE=$PWD/smap_b200/lib/libsmap_b200_base.so
timeout 300 python -m pytest tests/test_conv_gpu.py -x -q > gpurun_out/ab_pytest_conv.log 2>&1; echo "pytest conv rc=$?" > gpurun_out/ab_summary.txt
timeout 300 python tools/ab_hash.py > gpurun_out/ab_hash_new.txt 2> gpurun_out/ab_hash_new.err; echo "hash new rc=$?" >> gpurun_out/ab_summary.txt
SMAPB_LIB=$BASE timeout 300 python tools/ab_hash.py > gpurun_out/ab_hash_base.txt 2> gpurun_out/ab_hash_base.err; echo "hash base rc=$?" >> gpurun_out/ab_summary.txt
if cmp -s gpurun_out/ab_hash_new.txt gpurun_out/ab_hash_base.txt; then echo "digests IDENTICAL" >> gpurun_out/ab_summary.txt; else echo "digests DIFFER" >> gpurun_out/ab_summary.txt; fi
for rep in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --profile-csv gpurun_out/ab_ops_new.csv > gpurun_out/ab_bench_new_$rep.json 2> gpurun_out/ab_bench_new_$rep.err; echo "bench new $rep rc=$?" >> gpurun_out/ab_summary.txt
  SMAPB_LIB=$BASE timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --profile-csv gpurun_out/ab_ops_base.csv > gpurun_out/ab_bench_base_$rep.json 2> gpurun_out/ab_bench_base_$rep.err; echo "bench base $rep rc=$?" >> gpurun_out/ab_summary.txt
done
timeout 200 python tools/conv_micro.py > gpurun_out/ab_micro_new.txt 2>&1
SMAPB_LIB=$BASE timeout 200 python tools/conv_micro.py > gpurun_out/ab_micro_base.txt 2>&1
cat gpurun_out/ab_summary.txt; tail -3 gpurun_out/ab_pytest_conv.log
diff gpurun_out/ab_hash_new.txt gpurun_out/ab_hash_base.txt | head -20
for f in gpurun_out/ab_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f e2e %.1f ms %.3f clocks %s frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["frac"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
paste gpurun_out/ab_micro_new.txt gpurun_out/ab_micro_base.txt | cut -c1-160
