#!/bin/bash
# usage: tools/dev/ab.sh "ENV=a ..." "ENV=b ..."   -> the default bench line (20 steps after 5) under each environment, alternating, twice
cd /root/repo
B="timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra"
p() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['nan_envs'])"; }
for rep in 1 2; do for cfg in "$@"; do env $cfg $B 2>/dev/null | p "[$cfg]"; done; done
