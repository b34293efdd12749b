#!/bin/bash
# 8-GPU weak-scaling step with either collective (peer-memory kernel vs NCCL all-reduce)
mkdir -p gpurun_out
: > gpurun_out/scaling8.jsonl
NG=$(nvidia-smi -L | wc -l)
for c in ${COLLECTIVES:-p2p nccl p2p}; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $NG --steps 30 --warmup 5 --no-e2e --no-cpu --collective $c 2>/dev/null | grep '^{' | tail -1 >> gpurun_out/scaling8.jsonl
done
python - <<'PY'
import json
for line in open("gpurun_out/scaling8.jsonl"):
    d = json.loads(line)
    print(d["n_gpus"], "GPUs", d["config"]["parallelism"][-8:], round(d["value"] / 1e6, 1), "M frames/s", "ms/step", round(d["ms_per_step"], 4))
PY
