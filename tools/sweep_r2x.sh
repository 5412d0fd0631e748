#!/bin/bash
cd /root/repo
O=gpurun_out
run() { n=$1; shift; env "$@" 2>$O/sw_$n.err > $O/sw_$n.json
  python - <<PY
import json
try:
    d = json.loads(open("$O/sw_$n.json").read().strip().splitlines()[-1])
    print("$n", round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d.get("reward_stats_last_step",{}).get("mean_reward"))
except Exception as e:
    print("$n FAILED", e); print(open("$O/sw_$n.err").read()[-400:])
PY
}
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra"
run r2x_rg0 PBRL_REGROUP_AFTER=0 $B
run r2x_a4 PBRL_REGROUP_AFTER=4 $B
run r2x_a5 PBRL_REGROUP_AFTER=5 $B
run r2x_a2e3 PBRL_REGROUP_AFTER=2 PBRL_REGROUP_EVERY=3 $B
run r2x_a2e2 PBRL_REGROUP_AFTER=2 PBRL_REGROUP_EVERY=2 $B
run r2x_a3e3 PBRL_REGROUP_AFTER=3 PBRL_REGROUP_EVERY=3 $B
run r2x_a4_65k PBRL_REGROUP_AFTER=4 $B --envs 65536
run r2x_a0_65k PBRL_REGROUP_AFTER=0 $B --envs 65536
run r2x_a2e2_65k PBRL_REGROUP_AFTER=2 PBRL_REGROUP_EVERY=2 $B --envs 65536
