"""Compile the reference's model and motion DATA files into the assets this repo ships.

Run HERE (build container, /root/reference mounted).  Outputs are derived data, not copies:

  physics-based-rl_b200/assets/<model>.model.json   ModelConst (mjcf.py) of models/*.xml
  physics-based-rl_b200/assets/motions.npz          MotionTable (mocap.py) of motions/*.txt
  physics-based-rl_b200/assets/mocap_<name>.npz     data_pos / data_vel / durations per clip
  tests/golden/euler_rxyz.npz                       known answers of the reference's own
                                                    some_math/transformations.py (imported here)
  tests/golden/mocap_facts.json                     shapes/durations printed in the notebooks
"""
import importlib
import importlib.util
import json
import sys
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")
pkg = "physics-based-rl_b200"
mjcf = importlib.import_module(pkg + ".mjcf")
mocap = importlib.import_module(pkg + ".mocap")
ASSETS = ROOT / pkg / "assets"
GOLD = ROOT / "tests" / "golden"


def main():
    ASSETS.mkdir(exist_ok=True)
    for name in ("final_humanoid", "final_humanoid_no_gravity"):
        m = mjcf.compile_mjcf(REF / "models" / f"{name}.xml")
        (ASSETS / f"{name}.model.json").write_text(m.to_json())
        print(name, "nq", m.nq, "nv", m.nv, "mass", m.total_mass, "gravity", m.gravity)
    clips = mocap.load_motion_dir(REF / "motions")
    mocap.MotionTable(clips).save(ASSETS / "motions.npz")
    facts = {}
    for n, c in clips:
        np.savez_compressed(ASSETS / f"mocap_{n}.npz", data_pos=np.asarray(c.data_pos),
                            data_vel=np.asarray(c.data_vel), durations=c.durations,
                            total_time=c.total_time, loop=c.loop)
        facts[n] = dict(frames=len(c.data_pos), total_time=c.total_time, loop=c.loop,
                        dict1=c.duration_dict[1])
        print(n, facts[n])
    # notebook-printed facts: testing_forward_env.ipynb cells 6, 8, 26-31, 34, 41
    facts["_notebook"] = {
        "punch_shape": [65, 35], "getup_faceup_shape_pos": [227, 35], "getup_faceup_shape_vel": [227, 34],
        "getup_faceup_total_time": 3.7665161581999853,
        "punch_dict1": [0.03333330154418945, 0.0333333015],
        "getup_faceup_dict1": [0.016666000708937645, 0.0166660007]}
    (GOLD / "mocap_facts.json").write_text(json.dumps(facts, indent=1))

    warnings.simplefilter("ignore")
    spec = importlib.util.spec_from_file_location("ref_tr", REF / "some_math" / "transformations.py")
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    rng = np.random.default_rng(1234)
    q = rng.normal(size=(256, 4))
    q[0] = [0.06146124, 0, 0, 0.99810947]  # the reference's own doctest vector
    q[1] = [0, np.sqrt(0.5), 0, np.sqrt(0.5)]  # gimbal lock
    e = np.array([tr.euler_from_quaternion(v, axes="rxyz") for v in q])
    np.savez_compressed(GOLD / "euler_rxyz.npz", q_xyzw=q, euler=e)
    print("euler golden", e[0])


if __name__ == "__main__":
    main()
