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
python -m pytest tests -m gpu -x -q > $O/pytest_r2l.log 2>&1; tail -3 $O/pytest_r2l.log
run r2l_p16
run r2l_p48 PBRL_PHASE_LOCK=48
run r2l_p49 PBRL_PHASE_LOCK=49
run r2l_p52 PBRL_PHASE_LOCK=52
run r2l_nb_p16 PBRL_BENCH_NO_BOX_PAIRS=1
run r2l_nb_p48 PBRL_BENCH_NO_BOX_PAIRS=1 PBRL_PHASE_LOCK=48
PBRL_PHASE_LOCK=48 PBRL_LIB=/root/repo/scratch_libs/libpbrl_prof.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --stage-profile > $O/sw_prof_r2l.json 2> $O/sw_prof_r2l.err
