#!/bin/bash
# usage: tools/sweep.sh "<libs in scratch_libs>" "<phase-lock masks>" [workload]
cd /root/repo
for L in $1; do
  cp scratch_libs/$L physics-based-rl_b200/libpbrl.so
  for PL in $2; do
    export PBRL_PHASE_LOCK=$PL
    python bench.py --steps 16 --warmup 3 --workload ${3:-mimic_mixed} --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$L lock=$PL', round(d['value']), round(d['ms_per_step'],3))"
  done
done
