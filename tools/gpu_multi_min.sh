#!/bin/bash
# minimal 2-GPU check at HEAD: sharding-invariance tests under torchrun, then the bench at N = 2 (gathered path, NVML sampler per rank)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -x -s > gpurun_out/pytest_multigpu_n2.log 2>&1; echo "pytest multigpu rc=$?" > gpurun_out/summary_multi.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?" >> gpurun_out/summary_multi.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_of2.json 2> gpurun_out/bench_n1_of2.err; echo "bench n1 rc=$?" >> gpurun_out/summary_multi.txt
cat gpurun_out/summary_multi.txt; tail -4 gpurun_out/pytest_multigpu_n2.log; tail -c 900 gpurun_out/bench_n2.json | head -c 900; echo; python - <<'PY'
import json
for f in ("gpurun_out/bench_n1_of2.json","gpurun_out/bench_n2.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); print(f, d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], d["clocks"])
    except Exception as e: print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1200:])
PY
