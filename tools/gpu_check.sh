#!/bin/bash
# usage (on the GPU box): tools/gpu_check.sh <tag> -> GPU tests, smoke, default bench line
cd /root/repo
T=${1:-chk}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_$T.log 2>&1; tail -15 $O/pytest_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$T.log 2>&1; tail -1 $O/smoke_$T.log
python bench.py --steps 20 --warmup 5 > $O/bench_${T}_n1.json 2> $O/bench_${T}_n1.err; tail -3 $O/bench_${T}_n1.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_${T}_n1.json").read().strip().splitlines()[-1])
    print("n1", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "stationary", round(d["stationary"]["value"]), "cpu", d.get("cpu_baseline", {}).get("value"), d.get("clocks"))
    for k, v in d.get("extra_configs", {}).items():
        print("  ", k, v.get("value"), v.get("ms_per_step"), v.get("error"), (v.get("roofline") or {}).get("frac"))
except Exception as e:
    print("n1 FAILED", e)
PY
