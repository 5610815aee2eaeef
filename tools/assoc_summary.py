"""gpurun_out/assoc_ncu.csv + gpurun_out/assoc.log (tools/assoc_bw.py under ncu / with CUDA events) -> profiles/rNN_assoc_b64.txt"""
import collections
import csv
import json
import os
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else 6483.3
per = collections.OrderedDict()
for r in csv.reader(open("gpurun_out/assoc_ncu.csv", errors="replace")):
    if len(r) > 14 and r[0].isdigit():
        per.setdefault(r[4].split("(")[0], {})[r[12]] = float(r[14].replace(",", ""))
with open("profiles/%s_assoc_b64.txt" % tag, "w") as f:
    f.write("association kernels at B = 64 crowded scenes (15 persons per frame), B200\n")
    f.write("ncu --profile-from-start off -k regex:'nms|paf|group' --clock-control none --metrics gpu__time_duration.sum,"
            "dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed python tools/assoc_bw.py --ncu\n\n")
    for k, v in per.items():
        us = v["gpu__time_duration.sum"] / 1e3
        mb = (v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]) / 1e6
        gbs = mb / us * 1e3
        f.write("%-30s %6.1f us  %7.1f MB dram  %6.0f GB/s = %4.1f %% of the measured HBM peak (%.0f GB/s); ncu dram__throughput %4.1f %%\n"
                % (k.replace("void ", "")[:30], us, mb, gbs, 100 * gbs / peak, peak, v["dram__throughput.avg.pct_of_peak_sustained_elapsed"]))
    f.write("\nCUDA-event timing of the same calls (tools/assoc_bw.py, L2 flushed between repetitions):\n")
    f.write(open("gpurun_out/assoc.log").read())
    f.write("\nearlier: first half of round 2 paf_kernel 56.4 us (52.6 %), nms_flag 42.7 us (38.4 %); round 1 nms_kernel 221 us (5.8 %), "
            "paf_kernel 82 us (28.7 %), group_kernel 61 us\n")
print(open("profiles/%s_assoc_b64.txt" % tag).read())
