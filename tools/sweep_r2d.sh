#!/bin/bash
cd /root/repo
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "replay" 2>&1 | tail -2
for v in "" rr32 rr64; do
  if [ -z "$v" ]; then lib=/root/repo/physics-based-rl_b200/libpbrl.so; else lib=/root/repo/scratch_libs/libpbrl_$v.so; fi
  PBRL_LIB=$lib python bench.py --workload replay --steps 32 --warmup 4 > $O/sw_replay_$v.json 2> $O/sw_replay_$v.err
  python - <<PY
import json
d = json.loads(open("$O/sw_replay_$v.json").read().strip().splitlines()[-1])
print("replay $v", round(d["value"]/1e9,3), "G env-steps/s", round(d["ms_per_step"]*1e3,1), "us  frac", round(d["roofline"]["frac"],3), "per-step us", round(d["per_step_launch"]["ms"]*1e3,1), "e2e", round(d["e2e"]["value"]/1e9,2))
PY
done



