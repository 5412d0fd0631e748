#!/bin/bash
# dev sweep: CTA shapes and phase-lock masks of k_step (scratch libs built with -DPBRL_WARPS_PER_CTA=W -DPBRL_LOCK_GROUP=W)
cd /root/repo
O=gpurun_out
run() { # name lib phase
  PBRL_LIB=$2 PBRL_PHASE_LOCK=$3 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/sw_$1.json 2> $O/sw_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/sw_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$1 FAILED", e)
PY
}
D=/root/repo/physics-based-rl_b200/libpbrl.so
run w16_p16 $D 16
run w16_p0 $D 0
run w16_p17 $D 17
run w16_p24 $D 24
for w in 7 5 8; do
  run w${w}_p16 /root/repo/scratch_libs/libpbrl_w$w.so 16
  run w${w}_p0 /root/repo/scratch_libs/libpbrl_w$w.so 0
done
