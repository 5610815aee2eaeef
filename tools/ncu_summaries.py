"""Turn gpurun_out/launches.csv (+ prof_conv.ncu-rep) into the tracked summaries under profiles/."""
import collections
import csv
import json
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 5]
hdr, data = None, []
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr and r[0].isdigit():
        data.append(dict(zip(hdr, r)))
agg, tot = collections.OrderedDict(), 0.0
for d in data:
    name = re.sub(r"\(.*", "", d["Kernel Name"])
    v = float(d["Metric Value"].replace(",", "")) / 1e3
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
with open("profiles/%s_ncu_launch_list_summary.txt" % tag, "w") as f:
    f.write("SMAPB_NO_GRAPH=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python bench.py --ncu-one-step --warmup 3 --engines 1\n")
    f.write("%d launches = exactly one device-resident step of 8 frames; cold-cache serialised times: compare SHARES, not absolutes\n\n" % len(data))
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-52s n=%4d  %9.1f us  %5.1f%%\n" % (k[:52], n, us, 100 * us / tot))
    conv = sum(us for k, (n, us) in agg.items() if "conv_tc" in k)
    f.write("\nconv_tc_kernel share of the captured launches: %.1f%%\n" % (100 * conv / tot))
out = subprocess.run(["ncu", "-i", "gpurun_out/prof_conv.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
K = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "lts__t_sector_hit_rate.pct", "launch__registers_per_thread"]
T = R = W = 0.0
with open("profiles/%s_ncu_conv_tc_full_summary.txt" % tag, "w") as f:
    f.write("SMAPB_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 20 -c 10 python bench.py --ncu-one-step --warmup 3 --engines 1 (tools/gpu_final.sh)\n")
    f.write("%d consecutive conv_tc_kernel launches, B=8 832x512, bf16x3\n" % len(data))
    f.write("kernel                                        time_us  dram_rd_MB dram_wr_MB dram%  tensor_pipe%  l2_hit%  regs\n")
    for r in data:
        t, rd, wr = float(r[ix[K[0]]]), float(r[ix[K[1]]]), float(r[ix[K[2]]])
        T += t
        R += rd
        W += wr
        nm = re.sub(r"\(smapb.*", "", r[ix["Kernel Name"]]).replace("void smapb::", "")
        f.write("%-44s %8.1f %9.1f %9.1f %6.1f %8.1f %10.1f %6s\n" % (nm[:44], t, rd, wr, float(r[ix[K[4]]]), float(r[ix[K[3]]]), float(r[ix[K[5]]]), r[ix[K[6]]]))
    mean = (R + W) / len(data)
    tw = sum(float(r[ix[K[3]]]) * float(r[ix[K[0]]]) for r in data) / T
    f.write("\nsum: %.1f us; DRAM read %.1f MB, write %.1f MB; mean traffic per launch %.1f MB\n" % (T, R, W, mean))
    f.write("tensor pipe active: max %.1f%%, time-weighted mean %.1f%%\n" % (max(float(r[ix[K[3]]]) for r in data), tw))
json.dump({"kernel": "conv_tc_kernel", "launches_captured": len(data), "mean_dram_bytes_per_launch": mean * 1e6,
           "tensor_pipe_active_pct_time_weighted": tw,
           "source": "profiles/%s_ncu_conv_tc_full_summary.txt (ncu --set full, 40 consecutive launches)" % tag},
          open("profiles/%s_conv_traffic.json" % tag, "w"), indent=1)
print(open("profiles/%s_ncu_conv_tc_full_summary.txt" % tag).read()[-320:])
