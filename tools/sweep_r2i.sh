#!/bin/bash
cd /root/repo
O=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/plain_r2i.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 -f -o $O/prof_r2i_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_r2i_full.log 2>&1
PBRL_BENCH_NO_BOX_PAIRS=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/plain_r2i_nb.log 2>&1 &&
PBRL_BENCH_NO_BOX_PAIRS=1 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 -f -o $O/prof_r2i_nobox python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_r2i_nb.log 2>&1
ls -la $O/prof_r2i_*.ncu-rep
