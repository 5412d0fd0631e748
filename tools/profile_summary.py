#!/usr/bin/env python
"""Dev tool: turn the scratch outputs of tools/gpu_round.sh <tag> (gpurun_out/) into the committed summaries under profiles/.
usage: tools/profile_summary.py <tag> <name>     e.g.  tools/profile_summary.py r1b r1_final"""
import csv, io, json, subprocess, sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
G, P = ROOT / "gpurun_out", ROOT / "profiles"
tag, name = sys.argv[1], sys.argv[2]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]

rep = G / f"prof_{tag}.ncu-rep"
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
stalls = sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"))
with open(P / f"{name}_k_step_ncu_full_summary.csv", "w") as f:
    f.write(f"# PBRL_REGROUP_AFTER=0 (the whole env step in one launch; by default a large batch is two launches of 4 substeps each) ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1: python bench.py --steps 2 --warmup 3 --no-cpu-baseline ({rep.name})\n")
    f.write("metric,unit,value\n")
    for h in WANT + stalls:
        if h in hdr:
            i = hdr.index(h)
            f.write(f"{h},{units[i]},{vals[i]}\n")
get = lambda h: float(vals[hdr.index(h)].replace(",", ""))
scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
rd = get("dram__bytes_read.sum") * scale[units[hdr.index("dram__bytes_read.sum")]]
wr = get("dram__bytes_write.sum") * scale[units[hdr.index("dram__bytes_write.sum")]]
(P / f"{name.split(chr(95))[0]}_traffic.json").write_text(json.dumps({
    "kernel": "k_step", "workload": "mimic_mixed", "envs_per_gpu": 16384, "dram_bytes_per_launch": int(rd + wr),
    "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "algorithmic_bytes_per_launch": 16384 * 1248,
    "warp_instructions_per_launch": int(get("smsp__inst_executed.sum")), "warp_instructions_per_env_substep": get("smsp__inst_executed.sum") / (16384 * 8),
    "source": f"ncu --set full --clock-control none, one k_step launch (profiles/{name}_k_step_ncu_full_summary.csv)"}, indent=1))

# the replay-rollout kernel (HBM-bound)
rrep = G / f"prof_replay_{tag}.ncu-rep"
if rrep.exists():
    raw2 = subprocess.run(["ncu", "-i", str(rrep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r2 = list(csv.reader(io.StringIO(raw2)))
    h2, u2, v2 = r2[0], r2[1], r2[2]
    st2 = sorted(h for h in h2 if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"))
    with open(P / f"{name}_k_replay_gather_bulk_ncu_full_summary.csv", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:k_replay_gather_bulk -s 4 -c 1: python bench.py --workload replay --steps 4 --warmup 3 ({rrep.name})\n")
        f.write("metric,unit,value\n")
        for h in WANT + ["l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
                         "sm__cycles_active.min", "sm__cycles_active.max", "sm__cycles_elapsed.avg"] + st2:
            if h in h2:
                i = h2.index(h)
                f.write(f"{h},{u2[i]},{v2[i]}\n")
    g2 = lambda h: float(v2[h2.index(h)].replace(",", "")) * scale[u2[h2.index(h)]]
    tp = P / f"{name.split(chr(95))[0]}_traffic.json"
    tj = json.loads(tp.read_text())
    tj["k_replay_gather_bulk"] = {"kernel": "k_replay_gather_bulk", "workload": "replay", "envs_per_gpu": 4096, "replay_steps_per_launch": 256,
                             "dram_bytes_per_launch": int(g2("dram__bytes_read.sum") + g2("dram__bytes_write.sum")),
                             "dram_bytes_write": int(g2("dram__bytes_write.sum")), "algorithmic_bytes_per_launch": 4096 * 256 * 684,
                             "source": f"ncu --set full --clock-control none, one k_replay_gather_bulk launch (profiles/{name}_k_replay_gather_bulk_ncu_full_summary.csv)"}
    tp.write_text(json.dumps(tj, indent=1))

# launch list
lines = [l for l in (G / f"launches_{tag}.csv").read_text().splitlines() if l.startswith('"')]
(P / f"{name}_launches_ncu.csv").write_text("\n".join(lines) + "\n")
rd_ = csv.DictReader(io.StringIO("\n".join(lines)))
agg = OrderedDict()
for r in rd_:
    k = r["Kernel Name"].split("(")[0][:60]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += float(r["Metric Value"].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(P / f"{name}_launch_summary.csv", "w") as f:
    f.write("kernel,launches,total_ns,share\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f'"{k}",{v[0]},{int(v[1])},{v[1] / tot:.4f}\n')

# bench lines of the same call
with open(P / f"{name}_bench_lines.jsonl", "w") as f:
    for w in ("n1", "ref", "contact_free", "contact_rich", "cpu64", "replay", "nobalance", "noself"):
        p = G / f"bench_{tag}_{w}.json"
        if p.exists() and p.read_text().strip():
            f.write(p.read_text().strip().splitlines()[-1] + "\n")
print(open(P / f"{name}_k_step_ncu_full_summary.csv").read())
print(open(P / f"{name}_launch_summary.csv").read())
