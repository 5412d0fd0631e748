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
python -m pytest tests -m gpu -x -q > $O/pytest_r2m.log 2>&1; tail -3 $O/pytest_r2m.log
run r2m_full PBRL_PHASE_LOCK=48
run r2m_cvx0 PBRL_PHASE_LOCK=48 PBRL_LIB=/root/repo/scratch_libs/libpbrl_cvx0.so
run r2m_cvxcb PBRL_PHASE_LOCK=48 PBRL_LIB=/root/repo/scratch_libs/libpbrl_cvxcb.so
run r2m_cvxbb PBRL_PHASE_LOCK=48 PBRL_LIB=/root/repo/scratch_libs/libpbrl_cvxbb.so
run r2m_nb PBRL_PHASE_LOCK=48 PBRL_BENCH_NO_BOX_PAIRS=1
