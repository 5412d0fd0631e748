cd /root/repo
B="timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra"
p() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['nan_envs'])"; }
$B 2>/dev/null | p two_launch_r4
for r in 2 3 4 5; do PBRL_REGROUP_FUSED=1 PBRL_REGROUP_AFTER=$r $B 2>/dev/null | p fused_r$r; done
PBRL_REGROUP_FUSED=1 PBRL_REGROUP_AFTER=3 $B --envs 65536 2>/dev/null | p fused_r3_65k
