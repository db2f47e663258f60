"""torchrun worker of tests/test_multi_gpu.py: the sharded PPO update on W GPUs == the single-GPU update
(mobrob_b200/selfcheck.py: bit-identical parameters on every rank, both update paths)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from mobrob_b200.selfcheck import sharded_update_check


def main():
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    res = sharded_update_check(dev)
    res_car = sharded_update_check(dev, modes=("fused",), obs_dim=26, n_local=16, b_local=130)
    if dist.get_rank() == 0:
        print("DIST_RESULT " + json.dumps({"point": res, "car": res_car}) + f" ok={res['ok'] and res_car['ok']}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if res["ok"] and res_car["ok"] else 1)


if __name__ == "__main__":
    main()
