#!/bin/bash
# small single-box probes behind the numbers in DESIGN.md section 6 (each is one short gpurun call; none is a bench value):
#   roles     per-role wait cycles of every conv launch inside a real step          -> profiles/r02_roles_in_plan.txt
#   timeline  clock64 time line of CTA 0 (set-up, first operands, chunk ends, exit)  -> profiles/r02_conv_timeline_*.txt
#   halo      halo-strip variant against the generic pair tile on the 3x3 64->64 layer -> profiles/r02_halo_variant_ab.txt
#   toggles   bench with PDL / three handles against the default, back to back
#   retune    re-measured tile table for batch 8 against the committed one (bench under both, alternating)
mkdir -p gpurun_out
case "$1" in
  roles)
    SMAPB_ROLES_PLAN=gpurun_out/roles_plan.csv timeout 300 python tools/roles_plan.py | tee gpurun_out/roles_plan.txt ;;
  timeline)
    SMAPB_TIMELINE=1 timeout 120 python tools/conv_micro.py l3_c2 l3_c1 l3_c3 l4_c2 l2_c3 l1_c3 l1_c1 l1_c2 2>&1 | tee gpurun_out/timeline.txt
    SMAPB_FORCE_TILE=128,2 SMAPB_TIMELINE=1 timeout 120 python tools/conv_micro.py l3_c2 l3_c1 l4_c2 2>&1 | tee -a gpurun_out/timeline.txt ;;
  halo)
    timeout 300 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tee gpurun_out/halo.txt
    SMAPB_ROLES=1 SMAPB_TIMELINE=1 timeout 120 python tools/conv_micro.py l1_c2 2>&1 | tee -a gpurun_out/halo.txt
    SMAPB_FORCE_TILE=64,2 SMAPB_ROLES=1 timeout 120 python tools/conv_micro.py l1_c2 2>&1 | tee -a gpurun_out/halo.txt ;;
  toggles)
    b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/tg_$name.json 2> gpurun_out/tg_$name.err; }
    b default X=1; b pdl SMAPB_PDL=1; b engines3 SMAPB_BENCH_ENGINES=3; b default2 X=1
    python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/tg_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.1f e2e %.1f ms %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"]))
PY
    ;;
  retune)  # re-measure the tile table for the bench batch and compare the bench under both tables, alternating
    timeout 400 python tools/make_tile_table.py gpurun_out/b200_new.tsv 8 | tee gpurun_out/retune.txt
    diff <(grep -v "^#" smap_b200/tiles/b200.tsv | grep " 8x" | sort) <(grep -v "^#" gpurun_out/b200_new.tsv | sort) | tee -a gpurun_out/retune.txt
    b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/rt_$name.json 2> gpurun_out/rt_$name.err; }
    b old1 X=1; b new1 SMAPB_TILE_TABLE=$PWD/gpurun_out/b200_new.tsv; b old2 X=1; b new2 SMAPB_TILE_TABLE=$PWD/gpurun_out/b200_new.tsv
    python - <<'PY' | tee -a gpurun_out/retune.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/rt_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.1f e2e %.1f ms %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"]))
PY
    ;;
  *) echo "usage: $0 roles|timeline|halo|toggles|retune"; exit 2 ;;
esac
