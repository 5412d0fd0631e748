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
for r in 0 1 2 3 4; do run r2w_rg$r PBRL_REGROUP_AFTER=$r $B; done
run r2w_rg1_cr PBRL_REGROUP_AFTER=1 $B --workload contact_rich
run r2w_rg0_cr PBRL_REGROUP_AFTER=0 $B --workload contact_rich
