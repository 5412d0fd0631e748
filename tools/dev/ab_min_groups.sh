cd /root/repo
p() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['nan_envs'])"; }
for n in 4096 2368 8192; do for g in 2 1; do
PBRL_REGROUP_MIN_GROUPS=$g timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --envs $n 2>/dev/null | p "N=$n min_groups=$g"
done; done
