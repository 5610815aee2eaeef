#!/bin/bash
# end-of-round GPU pass in ONE box (round 2, second session): all -m gpu tests, smoke, bench (both arms), output digests against
# the committed ones, ncu launch list + per-launch DRAM / tensor-pipe metrics of one step, association at B = 64 (events + ncu),
# compute-sanitizer over the small-shape targets, one --set full capture of a few conv launches
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
timeout 1500 python -m pytest tests -q -m gpu -x -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $S
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 --profile-csv gpurun_out/ops.csv > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> $S
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> $S
timeout 300 python tools/ab_hash.py > gpurun_out/ab_hash_head.txt 2> gpurun_out/ab_hash_head.err; echo "digests rc=$?" >> $S
SMAPB_LIB=$PWD/smap_b200/lib/libsmap_b200_kymajor.so timeout 300 python tools/ab_hash.py > gpurun_out/ab_hash_kymajor.txt 2> gpurun_out/ab_hash_kymajor.err
if cmp -s gpurun_out/ab_hash_kymajor.txt profiles/r02_output_digests.txt; then echo "digests of the ky-major build IDENTICAL to profiles/r02_output_digests.txt (round-2 baseline build)" >> $S; else echo "digests of the ky-major build DIFFER from the baseline" >> $S; fi
SMAPB_NO_GRAPH=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --ncu-one-step --warmup 3 --engines 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu launch list rc=$?" >> $S
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
SMAPB_NO_GRAPH=1 timeout 600 ncu --profile-from-start off --clock-control none --metrics $METRICS --csv --log-file gpurun_out/conv_step.csv \
    python bench.py --ncu-one-step --warmup 3 --engines 1 > gpurun_out/ncu_step.log 2>&1; echo "ncu conv metrics rc=$?" >> $S
python tools/ncu_conv_summary.py gpurun_out/conv_step.csv gpurun_out/conv_traffic.json "ncu --profile-from-start off --metrics $METRICS python bench.py --ncu-one-step --engines 1 (one device-resident step of 8 frames, SMAPB_NO_GRAPH=1)" >> $S 2>&1
timeout 300 python tools/assoc_bw.py > gpurun_out/assoc.log 2>&1; echo "assoc rc=$?" >> $S
timeout 600 ncu --profile-from-start off -k regex:'nms|paf|group' --clock-control none \
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
   --csv --log-file gpurun_out/assoc_ncu.csv python tools/assoc_bw.py --ncu > gpurun_out/assoc_ncu.log 2>&1; echo "assoc ncu rc=$?" >> $S
rm -f gpurun_out/sanitizer_summary.txt
bash tools/gpu_sanitize.sh > /dev/null 2>&1; echo "sanitizer pass rc=$?" >> $S
rm -f gpurun_out/prof_conv.ncu-rep
SMAPB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 20 -c 10 -o gpurun_out/prof_conv \
    python bench.py --ncu-one-step --warmup 3 --engines 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $S
cat $S; grep -a "e2e \|passed\|failed" gpurun_out/pytest_gpu.log | tail -5; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_ref.json; cat gpurun_out/assoc.log; cat gpurun_out/sanitizer_summary.txt
