#!/bin/bash
# why is ms/step at N>1 above N=1?  (a) default (all-gather inside the graph) (b) all-gather outside the graph (c) NCCL limited to
# one CTA (d) no exchange at all: N independent replicas = the slowest GPU of the box
N=${1:-2}
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+RANDOM%200)) bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_$tag.json") if l.startswith("{")][-1]); print("$tag", "value=%.1f e2e=%.1f ms=%.3f per-rank=%s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["ms_per_step_per_rank"]))
except Exception as e: print("$tag FAILED", e, open("gpurun_out/scale_$tag.err").read()[-800:])
PY
}
for g in $(seq 0 $((N-1))); do CUDA_VISIBLE_DEVICES=$g timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('single gpu $g value=%.1f ms=%.3f'%(d['value'],d['ms_per_step']))"; done
run default A=1
run eager SMAPB_NCCL_EAGER=1
run maxctas1 NCCL_MAX_CTAS=1 NCCL_MIN_CTAS=1
run nogather SMAPB_BENCH_NO_GATHER=1
