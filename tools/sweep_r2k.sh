#!/bin/bash
# dev sweep r2k: default bench after the prefilter / sincos change, stage profile
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
python -m pytest tests -m gpu -x -q > $O/pytest_r2k.log 2>&1; tail -3 $O/pytest_r2k.log
run r2k_full
run r2k_noboxpairs PBRL_BENCH_NO_BOX_PAIRS=1
run r2k_noself PBRL_BENCH_NO_SELF_COLLISION=1
PBRL_LIB=/root/repo/scratch_libs/libpbrl_prof.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --stage-profile > $O/sw_prof_r2k.json 2> $O/sw_prof_r2k.err
head -5 $O/sw_prof_r2k.err
