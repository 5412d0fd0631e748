#!/bin/bash
cd /root/repo
O=gpurun_out
run() { n=$1; shift; env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/sw_$n.json 2> $O/sw_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/sw_$n.json").read().strip().splitlines()[-1])
    print("$n", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], d.get("nan_envs"))
except Exception as e:
    print("$n FAILED", e); print(open("$O/sw_$n.err").read()[-600:])
PY
}
run r2r_w16
run r2r_w17 PBRL_LIB=/root/repo/scratch_libs/libpbrl_w17.so
run r2r_w16b
run r2r_w17b PBRL_LIB=/root/repo/scratch_libs/libpbrl_w17.so
