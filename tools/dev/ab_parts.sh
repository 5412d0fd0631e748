cd /root/repo
p() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['nan_envs'])"; }
for n in 16384 4096; do for P in 2 3 4; do
PBRL_REGROUP_PARTS=$P timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --envs $n 2>/dev/null | p "N=$n parts=$P"
done; done
