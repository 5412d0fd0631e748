#!/bin/bash
cd /root/repo
O=gpurun_out
run() { n=$1; shift; env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/sw_$n.json 2> $O/sw_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/sw_$n.json").read().strip().splitlines()[-1])
    print("$n", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], d.get("self_contact_overflow_envs"))
except Exception as e:
    print("$n FAILED", e)
PY
}
run r2p_full
run r2p_cvx0 PBRL_LIB=/root/repo/scratch_libs/libpbrl_cvx0.so
run r2p_cvxcb PBRL_LIB=/root/repo/scratch_libs/libpbrl_cvxcb.so
run r2p_cvxbb PBRL_LIB=/root/repo/scratch_libs/libpbrl_cvxbb.so
PBRL_LIB=/root/repo/scratch_libs/libpbrl_prof.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --stage-profile > $O/sw_prof_r2p.json 2> $O/sw_prof_r2p.err
PBRL_LIB=/root/repo/scratch_libs/libpbrl_prof_cvx0.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --stage-profile > $O/sw_prof_r2p_cvx0.json 2> $O/sw_prof_r2p_cvx0.err
