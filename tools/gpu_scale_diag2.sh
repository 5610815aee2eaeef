#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+RANDOM%200)) bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_$tag.json") if l.startswith("{")][-1]); print("$tag", "value=%.1f e2e=%.1f ms=%.3f per-rank=%s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["ms_per_step_per_rank"]))
except Exception as e: print("$tag FAILED", e, open("gpurun_out/scale_$tag.err").read()[-800:])
PY
}

run async A=1
run syncgather SMAPB_BENCH_SYNC_GATHER=1


