#!/bin/bash
cd /root/repo
O=gpurun_out
run() { # name env...
  n=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/sw_$n.json 2> $O/sw_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/sw_$n.json").read().strip().splitlines()[-1])
    print("$n", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], d["nan_envs"], d["reward_stats_last_step"]["mean_reward"])
except Exception as e:
    print("$n FAILED", e)
PY
}
run base_p16 PBRL_PHASE_LOCK=16
for g in 8 4 2; do
  run vg${g}_p17 PBRL_LIB=/root/repo/scratch_libs/libpbrl_vg$g.so PBRL_PHASE_LOCK=17
  run vg${g}_p16 PBRL_LIB=/root/repo/scratch_libs/libpbrl_vg$g.so PBRL_PHASE_LOCK=16
done
