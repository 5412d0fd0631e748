#!/bin/bash
# usage (on the GPU box): tools/dev/gpu_bulk.sh <tag> -> replay tests (both store paths), replay bench of the shipped library and of the
# variant builds in scratch_libs/libpbrl_rb_*.so, the whole GPU suite, smoke, the default bench line, one ncu capture of the bulk kernel
cd /root/repo
T=${1:-r2bulk}
O=gpurun_out
S() { python - "$1" "$2" <<PY
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    sp = d.get("store_paths", {})
    print(sys.argv[1], "default", round(d["ms_per_step"], 4), "ms frac", round(d["roofline"]["frac"], 3), "| lsu", round(sp["ordinary_stores"]["ms_per_step"], 4),
          "bulk", round(sp["bulk_stores"]["ms_per_step"], 4), "per-item", round(d["per_item_kinematics"]["ms_per_step"], 4), "e2e", round(d["e2e"]["value"] / 1e9, 3), d.get("clocks"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
timeout 240 python -m pytest tests -m gpu -x -q -k replay_rollout > $O/pytest_${T}_replay.log 2>&1; tail -3 $O/pytest_${T}_replay.log
timeout 120 python bench.py --workload replay --steps 32 --warmup 4 > $O/bench_${T}_replay.json 2> $O/bench_${T}_replay.err; S shipped $O/bench_${T}_replay.json
timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest_$T.log 2>&1; tail -3 $O/pytest_$T.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$T.log 2>&1; tail -1 $O/smoke_$T.log
for L in scratch_libs/libpbrl_rb_*.so; do
  B=$(basename $L .so)
  PBRL_LIB=/root/repo/$L timeout 120 python bench.py --workload replay --steps 32 --warmup 4 > $O/bench_${T}_$B.json 2> $O/bench_${T}_$B.err; S $B $O/bench_${T}_$B.json
done
timeout 420 python bench.py --steps 20 --warmup 5 > $O/bench_${T}_n1.json 2> $O/bench_${T}_n1.err
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
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_replay_gather_bulk -s 4 -c 1 -f -o $O/prof_replay_$T python bench.py --workload replay --steps 4 --warmup 3 > $O/ncu_fr_$T.log 2>&1
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_replay_$T.csv python bench.py --workload replay --steps 4 --warmup 3 > $O/ncu_lr_$T.log 2>&1
ls -la $O | tail -4
