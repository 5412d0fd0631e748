#!/usr/bin/env python
"""Dev tool: per-device-function SASS statistics of one kernel in libpbrl.so.
usage: tools/sass_fn.py [lib.so] [kernel substring] [function substring to dump]"""
import re, subprocess, sys, tempfile, os, collections
lib = sys.argv[1] if len(sys.argv) > 1 else "physics-based-rl_b200/libpbrl.so"
kern = sys.argv[2] if len(sys.argv) > 2 else "k_step"
dump = sys.argv[3] if len(sys.argv) > 3 else None
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, check=True, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
lines = txt.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
cur, fns = "main", collections.OrderedDict()
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.match(r"^(\$?_Z\S+|\$__internal\S+):", l)
    if m:
        n = m.group(1)
        mm = re.search(r"\$_ZN4pbrl(\d+)([a-zA-Z_0-9]+)", n)
        cur = mm.group(2)[:int(mm.group(1))] if mm else ("main" if "$" not in n[1:] else n[:40])
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        fns.setdefault(cur, []).append(l)
tot = 0
for n, ins in fns.items():
    c = collections.Counter(re.sub(r"\..*", "", x.split("*/")[1].split()[0] if not x.split("*/")[1].split()[0].startswith("@") else x.split("*/")[1].split()[1]) for x in ins)
    tot += len(ins)
    print(f"{n:24s} {len(ins):6d} instr {len(ins)*16/1024:6.1f} KB  " + " ".join(f"{k}={v}" for k, v in c.most_common(9)))
print("total", tot, "instr", tot * 16 / 1024, "KB")
if dump:
    for n, ins in fns.items():
        if dump in n:
            print("\n".join(x[:120] for x in ins))
