"""Tiny mimic-mode run (for compute-sanitizer): 14 envs, 2 env steps, every entry point once."""
import importlib, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import ASSETS, load_model_asset, mocap, util_data, start_trajectory
envs = importlib.import_module("physics-based-rl_b200.envs")
m = load_model_asset("final_humanoid")
kp, kd = util_data.generate_kp_kd_gains()
env = envs.HumanoidTrain(None, None, 0.0, {}, m, kp, kd, num_envs=14, reward_mode="mimic", motion_table=mocap.MotionTable.load(ASSETS / "motions.npz"))
s = env.reset(1)
for k in range(2):
    s = env.step(s, s.pipeline_state.qpos[:, 7:].clone())
_ = s.pipeline_state.qM
rep = envs.HumanoidReplay(None, None, 0.0, {}, m, num_envs=14)
rs = rep.step(rep.reset(0), None, want_poses=True)
torch.cuda.synchronize()
print("sanity ok", float(s.reward[:, 0].mean()))
