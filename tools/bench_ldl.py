"""Dev: cycles per ldl_factor_solve at several occupancies (needs the -DPBRL_STAGE_PROF library via PBRL_LIB)."""
import ctypes, importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
L = importlib.import_module("physics-based-rl_b200._lib")
mjcf = importlib.import_module("physics-based-rl_b200.mjcf")
model = mjcf.load_model(ROOT / "physics-based-rl_b200" / "assets" / "final_humanoid.model.json")
M = L.Model(model, device=0)
lib = L.lib()
for w, c in [(1, 148), (4, 148), (8, 148), (8, 296), (1, 1)]:
    out = (ctypes.c_ulonglong * 3)()
    L.check(lib.pbrl_debug_bench_ldl(M.handle, w, c, 50, out))
    print(f"warps/cta {w} ctas {c}: avg {out[0] / max(out[2], 1):.0f} cycles, best {out[1]}")
