"""BASELINE.json configs[1]: batched evaluation of a pretrained policy over random goals
(examples/control.py --no-gui semantics: deterministic actions, terminate_on_goal, goal re-drawn on
reach; success = goal reached within `horizon` steps of being set).

  python tools/eval_policy.py [--zip tests/golden/policies/point-ppo.zip] [--n 16384] [--oracle-n 0]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def evaluate_gpu(zip_path, n, horizon=1000, seed=0, steps=None, env_name="point"):
    from mobrob_b200 import GpuVecEnv
    from mobrob_b200.ppo import PPO

    model = PPO.load(zip_path)
    env = GpuVecEnv(env_name, n, seed=seed, time_limit=horizon, terminate_on_goal=True)
    obs = env.reset_tensor()
    steps = steps or horizon
    first_len = torch.zeros(n, dtype=torch.int32, device=env.device)
    first_ok = torch.zeros(n, dtype=torch.bool, device=env.device)
    seen = torch.zeros(n, dtype=torch.bool, device=env.device)
    n_term = torch.zeros((), dtype=torch.int64, device=env.device)
    n_trunc = torch.zeros((), dtype=torch.int64, device=env.device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        act, _, _ = model.policy.forward_tensor(obs, None)
        obs, rew, done, trunc = env.step_tensor(act)
        d, tr = done.bool(), trunc.bool()
        new = d & ~seen
        first_len = torch.where(new, env.ep_len, first_len)
        first_ok = torch.where(new, ~tr, first_ok)
        seen |= d
        n_term += (d & ~tr).sum()
        n_trunc += tr.sum()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = first_ok & seen
    return {"n_envs": n, "steps": steps, "first_goal_success_rate": float(ok.float().mean()),
            "first_goal_mean_steps": float(first_len[ok].float().mean()),
            "all_goals_success_rate": float(n_term / torch.clamp(n_term + n_trunc, min=1)),
            "goals_reached": int(n_term), "timeouts": int(n_trunc),
            "env_steps_per_s": n * steps / dt,
            "first_len": first_len.cpu().numpy(), "first_ok": ok.cpu().numpy()}


def evaluate_oracle(zip_path, n, horizon=1000, seed=0, steps=None):
    import io
    import zipfile

    from oracle import point_oracle as po, sb3_oracle
    from oracle.vec_oracle import GoalVecOracle

    sd = torch.load(io.BytesIO(zipfile.ZipFile(zip_path).read("policy.pth")), map_location="cpu", weights_only=True)
    pol = sb3_oracle.MlpPolicyOracle(14)
    pol.load_state_dict(sd)
    env = GoalVecOracle(po.PointBody(n), seed=seed, time_limit=horizon, terminate_on_goal=True)
    obs = env.reset()
    first_len = np.zeros(n, np.int64)
    first_ok = np.zeros(n, bool)
    seen = np.zeros(n, bool)
    n_term = n_trunc = 0
    for _ in range(steps or horizon):
        with torch.no_grad():
            a = pol.predict_deterministic(torch.as_tensor(obs)).numpy()
        obs, rew, done, info = env.step(a)
        new = done & ~seen
        first_len[new] = info["ep_l"][new]
        first_ok[new] = ~info["truncated"][new]
        seen |= done
        n_term += int((done & ~info["truncated"]).sum())
        n_trunc += int(info["truncated"].sum())
    ok = first_ok & seen
    return {"n_envs": n, "first_goal_success_rate": float(ok.mean()),
            "first_goal_mean_steps": float(first_len[ok].mean()),
            "all_goals_success_rate": n_term / max(n_term + n_trunc, 1), "goals_reached": n_term, "timeouts": n_trunc,
            "first_len": first_len, "first_ok": ok}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--zip", default=os.path.join(ROOT, "tests", "golden", "policies", "point-ppo.zip"))
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--oracle-n", type=int, default=0)
    ap.add_argument("--env", default="point")
    a = ap.parse_args()
    if a.env != "point" and "point-ppo" in a.zip:
        a.zip = a.zip.replace("point-ppo", f"{a.env}-ppo")
    out = {"gpu": evaluate_gpu(a.zip, a.n, env_name=a.env)}
    if a.oracle_n:
        out["oracle"] = evaluate_oracle(a.zip, a.oracle_n)
        g = evaluate_gpu(a.zip, a.oracle_n)
        out["same_seeds"] = {"flags_equal": bool((g["first_ok"] == out["oracle"]["first_ok"]).all()),
                             "max_len_diff": int(np.abs(g["first_len"] - out["oracle"]["first_len"]).max())}
    for v in out.values():
        v.pop("first_len", None), v.pop("first_ok", None)
    print(json.dumps(out))
