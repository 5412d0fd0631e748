#!/bin/bash
cd /root/repo
O=gpurun_out
run() { n=$1; shift; env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra 2>$O/sw_$n.err > $O/sw_$n.json
  python - <<PY
import json
try:
    d = json.loads(open("$O/sw_$n.json").read().strip().splitlines()[-1])
    print("$n", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$n FAILED", e); print(open("$O/sw_$n.err").read()[-400:])
PY
}
run r2s_dw3
for dw in 1 8 16; do
run r2s_dw$dw PBRL_LIB=/root/repo/scratch_libs/libpbrl_dw$dw.so
run r2s_dw${dw}_sum PBRL_LIB=/root/repo/scratch_libs/libpbrl_dw$dw.so PBRL_COST_MODE=1
done
