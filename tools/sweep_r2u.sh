#!/bin/bash
cd /root/repo
O=gpurun_out
run() { n=$1; shift; env "$@" 2>$O/sw_$n.err > $O/sw_$n.json
  python - <<PY
import json
try:
    d = json.loads(open("$O/sw_$n.json").read().strip().splitlines()[-1])
    print("$n", round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["config"].get("envs_per_gpu"))
except Exception as e:
    print("$n FAILED", e); print(open("$O/sw_$n.err").read()[-400:])
PY
}
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra"
for ts in 1 0; do
run r2u_ts${ts}_16384 PBRL_TAIL_SPLIT=$ts $B
run r2u_ts${ts}_8192 PBRL_TAIL_SPLIT=$ts $B --envs 8192
run r2u_ts${ts}_4096 PBRL_TAIL_SPLIT=$ts $B --envs 4096
run r2u_ts${ts}_cpu64 PBRL_TAIL_SPLIT=$ts $B --workload cpu64
done
