#!/bin/bash
# multi-GPU pass (gpurun --gpus N -- 'bash tools/gpu_multi.sh N'): sharding-invariance test under torchrun, then bench at 1 and N GPUs
# on the same box (every GPU of the box is charged: N = 2 is enough to exercise the gathered path)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -x -s > gpurun_out/pytest_multigpu_n$N.log 2>&1; echo "pytest multigpu rc=$?" > gpurun_out/summary_multi.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_of$N.json 2> gpurun_out/bench_n1_of$N.err; echo "bench n1 rc=$?" >> gpurun_out/summary_multi.txt
for G in 2 4 8; do
  if [ $G -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29520+G)) bench.py --gpus $G --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$G.json 2> gpurun_out/bench_n$G.err; echo "bench n$G rc=$?" >> gpurun_out/summary_multi.txt
  fi
done
cat gpurun_out/summary_multi.txt; tail -5 gpurun_out/pytest_multigpu_n$N.log
python - <<'PY'
import json,glob
base=None
for f in sorted(glob.glob("gpurun_out/bench_n*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); n=d["n_gpus"]
        if n==1: base=d["value"]
        print(f, "n=%d value=%.1f e2e=%.1f ms/step=%.3f eff=%s"%(n,d["value"],d["e2e"]["value"],d["ms_per_step"], ("%.3f"%(d["value"]/n/base)) if base else "-"))
    except Exception as e:
        print(f,"FAILED",e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
