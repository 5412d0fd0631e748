#!/bin/bash
cd /root/repo
O=gpurun_out
run() { n=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $O/sw_$n.json 2> $O/sw_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/sw_$n.json").read().strip().splitlines()[-1])
    print("$n", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["config"]["self_collision"])
except Exception as e:
    print("$n FAILED", e, open("$O/sw_$n.err").read()[-400:])
PY
}
run noboxpairs PBRL_BENCH_NO_BOX_PAIRS=1
run noboxpairs_nomjx PBRL_BENCH_NO_BOX_PAIRS=1 PBRL_LIB=/root/repo/scratch_libs/libpbrl_nomjx.so
run full_nomjx PBRL_LIB=/root/repo/scratch_libs/libpbrl_nomjx.so
