"""Generate the golden fixtures in this directory from the reference checkout.

Run once in the build container (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``

Sources (all under /root/reference/data/policies/):
  * ``{point,car}-ppo.zip:data:_last_obs``   real MuJoCo 2.1.0 observations
  * ``{point,car}-ppo.zip:policy.pth``       SB3 2.0.0 MlpPolicy state dict
  * ``{point,car}-ppo.zip:policy.optimizer.pth``  Adam hyper-parameters
Outputs:
  * ``{env}_last_obs.npy``        float32 (n_envs, obs_dim)
  * ``{env}_policy.npz``          the 13 state-dict tensors, SB3 names
  * ``kat2.json``                 mu / V of the shipped policy on _last_obs (torch CPU fp32)
  * ``policies/{env}-ppo.zip``    byte copy of the shipped artefact (load/save tests)
  * ``ref_rng.json``              a few draws of the reference RNG streams (numpy)
  * ``examples/{train,control,fix_pickle_warning}.py``  byte copies of the reference's example scripts: INPUTS of
                                  tests/test_examples_gpu.py, which runs them unchanged through
                                  shims/ (they are fixtures, not product code -- nothing imports them)
  * ``kat4_ep_info.json``         the 100 most recent training episodes (r, l) stored in the shipped
                                  zips (``data:ep_info_buffer``, written by the real SB3 + MuJoCo run)
"""
import base64
import io
import json
import os
import pickle
import shutil
import sys
import warnings
import zipfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def forward(sd, obs):
    x = torch.as_tensor(obs)

    def mlp(prefix):
        h = torch.tanh(torch.nn.functional.linear(x, sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"]))
        return torch.tanh(torch.nn.functional.linear(h, sd[f"{prefix}.2.weight"], sd[f"{prefix}.2.bias"]))

    mu = torch.nn.functional.linear(mlp("mlp_extractor.policy_net"), sd["action_net.weight"], sd["action_net.bias"])
    v = torch.nn.functional.linear(mlp("mlp_extractor.value_net"), sd["value_net.weight"], sd["value_net.bias"])
    return mu.numpy(), v.numpy()[:, 0]


def main():
    kat2 = {}
    os.makedirs(os.path.join(HERE, "policies"), exist_ok=True)
    for env in ("point", "car"):
        src = f"{REF}/data/policies/{env}-ppo.zip"
        shutil.copyfile(src, os.path.join(HERE, "policies", f"{env}-ppo.zip"))
        z = zipfile.ZipFile(src)
        data = json.loads(z.read("data"))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            last_obs = pickle.loads(base64.b64decode(data["_last_obs"][":serialized:"]))
        last_obs = np.asarray(last_obs, dtype=np.float32)
        np.save(os.path.join(HERE, f"{env}_last_obs.npy"), last_obs)
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
        np.savez(os.path.join(HERE, f"{env}_policy.npz"), **{k: v.numpy() for k, v in sd.items()})
        mu, v = forward(sd, last_obs)
        kat2[env] = {"mu": mu.astype(np.float64).tolist(), "v": v.astype(np.float64).tolist()}
        print(env, last_obs.shape, mu, v)
    with open(os.path.join(HERE, "kat2.json"), "w") as f:
        json.dump(kat2, f, indent=1)

    # the example scripts, byte for byte
    os.makedirs(os.path.join(HERE, "examples"), exist_ok=True)
    for name in ("train.py", "control.py", "fix_pickle_warning.py"):
        shutil.copyfile(f"{REF}/examples/{name}", os.path.join(HERE, "examples", name))

    # KAT-4: Monitor's episode records of the reference's own training run
    kat4 = {}
    for env in ("point", "car"):
        data = json.loads(zipfile.ZipFile(f"{REF}/data/policies/{env}-ppo.zip").read("data"))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            eps = pickle.loads(base64.b64decode(data["ep_info_buffer"][":serialized:"]))
        kat4[env] = {"r": [float(e["r"]) for e in eps], "l": [int(e["l"]) for e in eps],
                     "n_steps": data["n_steps"], "n_envs": data["n_envs"], "num_timesteps": data["num_timesteps"]}
        print(env, "ep_info_buffer:", len(eps), "episodes, mean length", np.mean(kat4[env]["l"]))
    with open(os.path.join(HERE, "kat4_ep_info.json"), "w") as f:
        json.dump(kat4, f)

    # Reference RNG streams (numpy is the arithmetic; this freezes our reading of the call order)
    from oracle import ref_rng

    rng = {"heading": {str(s): ref_rng.engine_heading(s) for s in (1, 2, 3, 7, 1000, 65537)}}
    b = ref_rng.init_box()
    b.seed(0)
    rng["init_seed0"] = [b.sample().astype(np.float64).tolist() for _ in range(3)]
    g = ref_rng.goal_box()
    g.seed(1)
    rng["goal_seed1"] = [g.sample().astype(np.float64).tolist() for _ in range(3)]
    with open(os.path.join(HERE, "ref_rng.json"), "w") as f:
        json.dump(rng, f, indent=1)


if __name__ == "__main__":
    main()
