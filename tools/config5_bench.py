"""BASELINE.json configs[4]: 1024x1024 input (256x256 maps), 1 GPU - backbone-only microbench plus the whole path (the
association runs at any map size since round 2).  CUDA-event timing after warm-up; `--ncu` brackets ONE eager backbone forward
with cudaProfilerStart/Stop for `ncu --profile-from-start off`.
    python tools/config5_bench.py [--batch 8] [--ncu]  ->  one JSON line per mode"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from smap_b200 import schema
from smap_b200.engine import RECORD_BYTES, Engine, scale_row

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--ncu", action="store_true")
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
B, H, W = a.batch, 1024, 1024
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {}
peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
eng = Engine(0, max_batch=B, in_h=H, in_w=W, stream=torch.cuda.Stream())
eng.load_state_dict(schema.make_state_dict(0, "identity"))
xs = [schema.make_input(B, H, W, seed=1 + r).cuda() for r in range(2)]
sc = dict(scale=0.5, img_width=2048, img_height=2048, net_width=W, net_height=H, f_x=2048.0, f_y=2048.0, cx=1024.0, cy=1024.0)
scales = torch.from_numpy(np.stack([scale_row(sc)] * B)).cuda()
out = torch.empty(B, RECORD_BYTES, dtype=torch.uint8, device="cuda")
n_conv, flops = eng.plan_info(B)
torch.cuda.synchronize()
if a.ncu:
    with torch.cuda.stream(eng.stream):
        eng.forward(xs[0])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    with torch.cuda.stream(eng.stream):
        eng.forward(xs[1])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    eng.close()
    sys.exit(0)


def timed(fn, steps):
    cur = torch.cuda.current_stream()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    eng.stream.wait_event(e0)
    for i in range(steps):
        fn(i)
    cur.wait_stream(eng.stream)
    e1.record(cur)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


hm = torch.empty(B, 43, H // 4, W // 4, device="cuda")
for i in range(4):
    eng.infer_device(xs[i % 2], scales, out=out)
torch.cuda.synchronize()
ms_path = timed(lambda i: eng.infer_device(xs[i % 2], scales, out=out), a.steps)
with torch.cuda.stream(eng.stream):
    for i in range(2):
        eng.forward(xs[i % 2])
torch.cuda.synchronize()


def fwd(i):
    with torch.cuda.stream(eng.stream):
        eng.forward(xs[i % 2])


ms_bb = timed(fwd, a.steps)
eng.profile_begin()
for i in range(3):
    eng.infer_device(xs[i % 2], scales, out=out)
prof = eng.profile_end(None)
conv_ms = prof["conv"][0] / 3
tot_ms = sum(v[0] for v in prof.values()) / 3
line = {"workload": "configs[4]: %dx%d input, batch %d, 1 GPU" % (W, H, B), "conv_launches": n_conv,
        "algorithmic_gflop_per_frame": flops / B * 1e-9,
        "whole_path_graph_replay": {"ms_per_batch": ms_path, "frames_per_s": B / ms_path * 1e3},
        "backbone_only_eager_c_abi": {"ms_per_batch": ms_bb, "frames_per_s": B / ms_bb * 1e3},
        "conv_kernel": {"ms_per_batch_serialised_events": conv_ms, "share": conv_ms / tot_ms,
                        "achieved_tflops_algorithmic_in_graph_mode": flops / (conv_ms / tot_ms * ms_path * 1e-3) * 1e-12,
                        "frac_of_measured_bf16_sustained": flops / (conv_ms / tot_ms * ms_path * 1e-3) * 1e-12 / peak_tf,
                        "tensor_pipe_flop_multiplier": 3},
        "gpu": torch.cuda.get_device_name(0)}
print(json.dumps(line))
eng.close()
