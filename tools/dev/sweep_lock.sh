#!/bin/bash
cd /root/repo
for PL in $1; do
  export PBRL_PHASE_LOCK=$PL
  python bench.py --steps 16 --warmup 3 --workload ${2:-mimic_mixed} --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lock=$PL', round(d['value']), round(d['ms_per_step'],3))"
done
