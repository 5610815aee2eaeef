#!/bin/bash
# ncu evidence for profiles/: per-launch device time + DRAM bytes + tensor-pipe activity of exactly one bench step (conv traffic
# summary the bench line quotes), the same for config 5 (1024x1024, batch 8), and a --set full capture of representative conv launches.
# Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
SMAPB_NO_GRAPH=1 timeout 900 ncu --profile-from-start off --clock-control none --metrics $METRICS --csv --log-file gpurun_out/conv_step.csv \
    python bench.py --ncu-one-step --warmup 3 --engines 1 > gpurun_out/ncu_step.log 2>&1
python tools/ncu_conv_summary.py gpurun_out/conv_step.csv gpurun_out/conv_traffic.json "ncu --profile-from-start off --metrics $METRICS python bench.py --ncu-one-step --engines 1 (one device-resident step of 8 frames, SMAPB_NO_GRAPH=1)"
for B in 1 8; do timeout 600 python tools/config5_bench.py --batch $B >> gpurun_out/config5_bench.jsonl; done
SMAPB_NO_GRAPH=1 timeout 900 ncu --profile-from-start off --clock-control none --metrics $METRICS --csv --log-file gpurun_out/config5_ncu_b8.csv \
    python tools/config5_bench.py --batch 8 --ncu > gpurun_out/config5_ncu.log 2>&1
python tools/ncu_conv_summary.py gpurun_out/config5_ncu_b8.csv gpurun_out/config5_conv_ncu_b8.json "one eager backbone forward, 1024x1024, batch 8"
SMAPB_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 60 -c 24 -o gpurun_out/prof_conv \
    python bench.py --ncu-one-step --warmup 3 --engines 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
