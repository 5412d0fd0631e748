#!/bin/bash
# dev sweep: work-estimate variants of the balanced order (PBRL_COST_MODE), stage profile with the vote counters
cd /root/repo
O=gpurun_out
run() { # name env...
  n=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/sw_$n.json 2> $O/sw_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/sw_$n.json").read().strip().splitlines()[-1])
    print("$n", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$n FAILED", e)
PY
}
run s6_cost0 PBRL_COST_MODE=0
run s6_cost1 PBRL_COST_MODE=1
run s6_cost2 PBRL_COST_MODE=2
run s6_nobal PBRL_BENCH_NO_BALANCE=1
for c in 0 1 2; do
PBRL_COST_MODE=$c PBRL_LIB=/root/repo/scratch_libs/libpbrl_prof.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --stage-profile > $O/sw_prof$c.json 2> $O/sw_prof$c.err
grep -E "votes|cycles per env|iterations per" $O/sw_prof$c.err
done
