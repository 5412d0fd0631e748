#!/bin/bash
# usage (on a GPU box with N GPUs): tools/gpu_scale.sh <tag> <N>  -> weak (16 384 envs per GPU) and strong (65 536 envs in total) scaling lines
cd /root/repo
T=${1:-r2}; N=${2:-2}; O=gpurun_out
for S in weak strong; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --scaling $S --no-cpu-baseline --no-extra > $O/scale_${T}_${S}_n$N.json 2> $O/scale_${T}_${S}_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/scale_${T}_${S}_n$N.json").read().strip().splitlines()[-1])
    print("$S n=$N", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["scaling"], d["config"].get("envs_per_gpu"), d["clocks"])
except Exception as e:
    print("$S n=$N FAILED", e); print(open("$O/scale_${T}_${S}_n$N.err").read()[-800:])
PY
done
