#!/bin/bash
cd /root/repo
O=gpurun_out
run() { n=$1; shift; env "$@" 2>$O/sw_$n.err > $O/sw_$n.json
  python - <<PY
import json
try:
    d = json.loads(open("$O/sw_$n.json").read().strip().splitlines()[-1])
    print("$n", round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$n FAILED", e); print(open("$O/sw_$n.err").read()[-400:])
PY
}
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra"
run r2v_base $B
run r2v_r96 PBRL_LIB=/root/repo/scratch_libs/libpbrl_r96.so $B
run r2v_r80 PBRL_LIB=/root/repo/scratch_libs/libpbrl_r80.so $B
