#!/bin/bash
# usage (on the GPU box): tools/gpu_check.sh [tag]   -> parity tests, a short bench, the stage profile
cd /root/repo
TAG=${1:-cur}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/b_$TAG.json 2> gpurun_out/b_$TAG.err
python -c "import sys,json; d=json.loads(open('gpurun_out/b_$TAG.json').read()); print('bench', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'nan', d['nan_envs'], d['reward_stats_last_step'])"
if [ -f scratch_libs/libpbrl_prof.so ]; then
  PBRL_LIB=/root/repo/scratch_libs/libpbrl_prof.so python bench.py --steps 8 --warmup 3 --no-cpu-baseline --stage-profile 2> gpurun_out/stage_$TAG.txt > /dev/null
  cat gpurun_out/stage_$TAG.txt
fi
