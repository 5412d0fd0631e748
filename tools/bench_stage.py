"""Dev: warm-cache cycles per call of each stage function at 1 and 16 warps per SM (needs the -DPBRL_STAGE_PROF
library via PBRL_LIB).  usage: PBRL_LIB=scratch_libs/libpbrl_prof.so python tools/bench_ldl.py [stage ids...]"""
import ctypes, importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
L = importlib.import_module("physics-based-rl_b200._lib")
mjcf = importlib.import_module("physics-based-rl_b200.mjcf")
model = mjcf.load_model(ROOT / "physics-based-rl_b200" / "assets" / "final_humanoid.model.json")
M = L.Model(model, device=0)
lib = L.lib()
NAMES = ["ldl_factor_solve", "stage_fk", "stage_com", "stage_crb", "stage_collide", "stage_comvel", "stage_bias", "stage_rows",
         "tree_Mx", "-", "rows_JTf", "assemble_hessian", "rows_Jx", "dense_factor_solve", "stage_collide_self"]
ids = [int(a) for a in sys.argv[1:]] or range(len(NAMES))
for s in ids:
    r = []
    for w, c in [(1, 148), (16, 148)]:
        out = (ctypes.c_ulonglong * 3)()
        L.check(lib.pbrl_debug_bench_stage(M.handle, s, w, c, 50, out))
        r.append(f"{w * c // 148:2d} warps/SM: avg {out[0] / max(out[2], 1):7.0f} best {out[1]:6d}")
    print(f"{NAMES[s]:18s} " + "   ".join(r))
