#!/bin/bash
# weak-scaling bench at 1, 2, 4, 8 GPUs of one box (as many as the box has); one JSON line per N
mkdir -p gpurun_out
: > gpurun_out/scaling.jsonl
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  [ $n -gt $NG ] && break
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-e2e --no-cpu 2>/dev/null | tail -1 >> gpurun_out/scaling.jsonl
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $n --steps 20 --warmup 5 --no-e2e --no-cpu 2>/dev/null | grep '^{' | tail -1 >> gpurun_out/scaling.jsonl
  fi
done
python - <<'PY'
import json
base = None
for line in open("gpurun_out/scaling.jsonl"):
    d = json.loads(line)
    base = base or d["value"]
    print(d["n_gpus"], "GPUs", round(d["value"] / 1e6, 1), "M frames/s", "ms/step", round(d["ms_per_step"], 4),
          "efficiency", round(d["value"] / (base * d["n_gpus"]), 3))
PY
