#!/bin/bash
# usage (on the GPU box): tools/gpu_round.sh <tag>  -> GPU tests, the bench lines of every workload, the reference arm, ncu launch list + one full capture
cd /root/repo
T=${1:-r1}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_$T.log 2>&1; tail -3 $O/pytest_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$T.log 2>&1; tail -1 $O/smoke_$T.log
python bench.py --steps 20 --warmup 5 > $O/bench_${T}_n1.json 2> $O/bench_${T}_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_${T}_ref.json 2> $O/bench_${T}_ref.err
for W in contact_free contact_rich cpu64; do python bench.py --workload $W --steps 16 --warmup 3 --no-cpu-baseline > $O/bench_${T}_$W.json 2>> $O/bench_${T}_n1.err; done
python bench.py --workload replay --steps 64 --warmup 4 > $O/bench_${T}_replay.json 2>> $O/bench_${T}_n1.err
PBRL_BENCH_NO_BALANCE=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${T}_nobalance.json 2>> $O/bench_${T}_n1.err
PBRL_BENCH_NO_SELF_COLLISION=1 python bench.py --steps 20 --warmup 5 > $O/bench_${T}_noself.json 2>> $O/bench_${T}_n1.err
for f in n1 ref contact_free contact_rich cpu64 replay nobalance noself; do python - <<PY
import json
try:
    d = json.loads(open("$O/bench_${T}_$f.json").read())
    print("$f", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d.get("e2e", {}).get("value", 0)), "cpu", d.get("cpu_baseline", {}).get("value"), d.get("clocks"))
except Exception as e:
    print("$f FAILED", e)
PY
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain_$T.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_l_$T.log 2>&1
PBRL_REGROUP_AFTER=0 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 -f -o $O/prof_$T python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_f_$T.log 2>&1
python bench.py --workload replay --steps 4 --warmup 3 > $O/plain_replay_$T.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_replay_gather_bulk -s 4 -c 1 -f -o $O/prof_replay_$T python bench.py --workload replay --steps 4 --warmup 3 > $O/ncu_fr_$T.log 2>&1
ls -la $O | tail -5
