#!/usr/bin/env python
"""Dev tool: aggregate an ncu --set full --import-source report of k_step per device function.
usage: tools/ncu_fn.py report.ncu-rep lib.so [envs substeps_per_launch]   (lib.so = the library the report was taken with)"""
import bisect, csv, io, os, re, subprocess, sys, tempfile
rep, lib = sys.argv[1], sys.argv[2]
envs = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
nf = int(sys.argv[4]) if len(sys.argv) > 4 else 8
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, check=True, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    lines = subprocess.run(["nvdisasm", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and "k_step" in l][0]
labels = []
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.match(r"^(\$?_Z\S+|\$__internal\S+):", l)
    if m:
        n = m.group(1)
        mm = re.search(r"\$_ZN4pbrl(\d+)([a-zA-Z_0-9]+)", n)
        labels.append([mm.group(2)[:int(mm.group(1))] if mm else ("main" if "$" not in n[1:] else n[:30]), None])
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m and labels and labels[-1][1] is None:
        labels[-1][1] = int(m.group(1), 16)
labels = [(n, o) for n, o in labels if o is not None]
offs = [o for _, o in labels]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr, data = rows[hi], rows[hi + 1:]
ia, iex, ith, ismp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
iwf, iwx = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][ia], 16)
agg = {}
for r in data:
    if len(r) < len(hdr):
        continue
    k = bisect.bisect_right(offs, int(r[ia], 16) - base) - 1
    a = agg.setdefault(labels[k][0], [0, 0, 0, 0, {}, 0, 0])
    a[0] += 1; a[1] += int(r[iex]); a[2] += int(r[ith]); a[3] += int(r[ismp]); a[5] += int(r[iwf]); a[6] += int(r[iwx])
    for i in stall:
        a[4][hdr[i]] = a[4].get(hdr[i], 0) + int(r[i])
tex, ts = sum(a[1] for a in agg.values()), sum(a[3] for a in agg.values())
print("total warp inst %.3e  per env-substep %.0f" % (tex, tex / (envs * nf)))
twf = sum(a[5] for a in agg.values())
print("shared-memory wavefronts per env-substep %.0f (excessive %.0f)" % (twf / (envs * nf), sum(a[6] for a in agg.values()) / (envs * nf)))
print(f"{'function':20s} {'sass':>6s} {'exec%':>6s} {'inst/substep':>12s} {'thr/inst':>8s} {'time%':>6s} {'cyc/inst(rel)':>8s} {'smem wf/substep':>15s} {'excess':>7s}  top stalls (share of the function's samples)")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][3]):
    if a[1] == 0:
        continue
    top = sorted(a[4].items(), key=lambda kv: -kv[1])[:6]
    print(f"{n:20s} {a[0]:6d} {100*a[1]/tex:6.2f} {a[1]/(envs*nf):12.0f} {a[2]/max(a[1],1):8.1f} {100*a[3]/ts:6.2f} {a[3]/max(a[1],1)*tex/ts:8.2f} {a[5]/(envs*nf):15.0f} {a[6]/(envs*nf):7.0f}  "
          + " ".join(f"{k[6:]}={v/max(a[3],1):.2f}" for k, v in top))
