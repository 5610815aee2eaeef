#!/bin/bash
# cheap switches on the bench at HEAD (one box, back to back): default, PDL, three handles, default again
mkdir -p gpurun_out
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/tg_$name.json 2> gpurun_out/tg_$name.err; }
b default X=1
b pdl SMAPB_PDL=1
b engines3 SMAPB_BENCH_ENGINES=3
b default2 X=1
timeout 300 python tools/ab_hash.py > gpurun_out/ab_hash_head.txt 2>&1
for f in gpurun_out/tg_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f e2e %.1f ms %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
diff gpurun_out/ab_hash_head.txt gpurun_out/ab_hash_base.txt && echo "digests identical to the round-2 baseline build"
