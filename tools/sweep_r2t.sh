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
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_r2t.log 2>&1; tail -3 $O/pytest_r2t.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra"
run r2t_16384 $B
run r2t_8192 $B --envs 8192
run r2t_4096 $B --envs 4096
run r2t_cpu64 $B --workload cpu64
run r2t_65536 $B --envs 65536
