#!/bin/bash
cd /root/repo
O=gpurun_out
run() { n=$1; shift; env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/sw_$n.json 2> $O/sw_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/sw_$n.json").read().strip().splitlines()[-1])
    print("$n", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$n FAILED", e)
PY
}
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_r2o.log 2>&1; tail -5 $O/pytest_r2o.log
run r2o_p48
run r2o_p16 PBRL_PHASE_LOCK=16
run r2o_nb PBRL_BENCH_NO_BOX_PAIRS=1
PBRL_LIB=/root/repo/scratch_libs/libpbrl_prof.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --stage-profile > $O/sw_prof_r2o.json 2> $O/sw_prof_r2o.err
grep -E "cycles per|collide|convex" $O/sw_prof_r2o.err
