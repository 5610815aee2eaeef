#!/bin/bash
# ncu evidence for profiles/: (1) launch list with per-launch device time for one bench step, (2) --set full capture of the
# dominant kernel (conv_tc) on representative launches.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
export SMAPB_NO_GRAPH=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 460 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 230 -c 40 -o gpurun_out/prof_conv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
timeout 900 ncu --profile-from-start off --clock-control none -k regex:conv_tc \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --csv --log-file gpurun_out/conv_step.csv python bench.py --ncu-one-step --warmup 3 > gpurun_out/ncu_step.log 2>&1
echo "one-step conv metrics rc=$?"
ls -la gpurun_out | tail -8
