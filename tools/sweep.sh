#!/bin/bash
# usage: tools/sweep.sh "<lib list>" : runs bench for each scratch lib with and without the phase lock
cd /root/repo/physics-based-rl_b200
for L in $1; do
  cp ../scratch_libs/$L libpbrl.so
  for PL in 0 1; do
    if [ $PL = 0 ]; then export PBRL_NO_PHASE_LOCK=1; else unset PBRL_NO_PHASE_LOCK; fi
    python ../bench.py --steps 16 --warmup 3 --workload ${2:-mimic_mixed} --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$L lock=$PL', round(d['value']), round(d['ms_per_step'],3))"
  done
done
