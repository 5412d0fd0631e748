#!/usr/bin/env python
"""Dev tool: SASS instruction count per source line of one device function of k_step (needs -lineinfo).
usage: tools/sass_lines.py <function substring> [lib.so] [top N]"""
import re, subprocess, tempfile, os, collections, sys
fn = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "physics-based-rl_b200/libpbrl.so"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, check=True, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-c", "-g", os.path.join(td, cub)], capture_output=True, text=True).stdout
lines = txt.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and "k_step" in l][0]
infn, cur, per = False, None, collections.Counter()
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.match(r"^(\$?_Z\S+|\$__internal\S+):", l)
    if m:
        infn = fn in m.group(1)
        continue
    mm = re.search(r'//## File "(.*?)", line (\d+)', l)
    if mm:
        cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
        continue
    if infn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        per[cur] += 1
srcs = {}
for (f, ln), c in sorted(per.items(), key=lambda kv: -kv[1])[:top]:
    if f not in srcs:
        p = os.path.join("physics-based-rl_b200/csrc", f)
        srcs[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = srcs[f][ln - 1].strip()[:100] if ln - 1 < len(srcs[f]) else ""
    print(f"{c:5d} {f}:{ln}  {text}")
print("total", sum(per.values()))
