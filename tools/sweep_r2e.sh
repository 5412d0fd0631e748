#!/bin/bash
cd /root/repo
O=gpurun_out
PBRL_LIB=/root/repo/scratch_libs/libpbrl_prof.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --stage-profile > $O/sw_prof_r2e.json 2> $O/sw_prof_r2e.err
cat $O/sw_prof_r2e.err | head -50
PBRL_BENCH_NO_SELF_COLLISION=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $O/sw_noself_r2e.json 2>/dev/null
python - <<PY
import json
d = json.loads(open("$O/sw_noself_r2e.json").read().strip().splitlines()[-1])
print("noself", round(d["value"]), round(d["ms_per_step"], 3))
PY
