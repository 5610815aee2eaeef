#!/bin/bash
# association: new streaming NMS, GT lift, runtime map sizes; B=64 timing + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_assoc_gpu.py tests/test_f4_gpu.py tests/test_pipeline_gpu.py tests/test_shims_gpu.py tests/test_e2e_chain_gpu.py tests/test_ref_driver_gpu.py -q -m gpu -x -s > gpurun_out/pytest_assoc.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary.txt
timeout 300 python tools/assoc_bw.py > gpurun_out/assoc.log 2>&1; echo "assoc rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --profile-from-start off -k regex:'nms|paf|group' --clock-control none \
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
   --csv --log-file gpurun_out/assoc_ncu.csv python tools/assoc_bw.py --ncu > gpurun_out/assoc_ncu.log 2>&1; echo "assoc ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -a "e2e \|passed\|failed\|Error\|assert" gpurun_out/pytest_assoc.log | tail -12; cat gpurun_out/assoc.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/assoc_ncu.csv")) if len(r)>14 and r[0].isdigit()]
agg={}
for r in rows: agg.setdefault((r[0],r[4][:40]),{})[r[12]]=float(r[14].replace(",",""))
for (i,k),m in agg.items():
    t=m["gpu__time_duration.sum"]; b=m["dram__bytes_read.sum"]+m["dram__bytes_write.sum"]
    print(k, "%.1f us"%(t/1e3), "%.1f MB"%(b/1e6), "%.0f GB/s"%(b/t), "ncu dram%%=%.1f"%m["dram__throughput.avg.pct_of_peak_sustained_elapsed"])
PY
