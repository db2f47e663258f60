#!/usr/bin/env python
"""bench.py -- env-steps/sec (rollout + PPO update), point env, N x B200.

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA hot path)
  python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle port of the
                                                             # reference stack on the host cores

A "step" is one PPO iteration of the workload in data/configs/point-ppo-b200.yaml on every
rank: fused rollout of 296 steps x 4096 envs (1 212 416 env-steps), GAE, then 10 epochs x 64
minibatches of 18 944 samples (forward, clipped loss, backward, grad-norm clip, Adam).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions used.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ENVS = 4096          # per GPU (weak scaling)
N_STEPS = 296
BATCH = 18944
N_EPOCHS = 10
METRIC = "env-steps/sec (rollout+PPO update), point env"
WORKLOAD = "point env 4096 parallel envs/GPU full rollout + GAE + PPO update (BASELINE.json configs[2])"

# algorithmic work of the dominant kernel (ppo_epoch_tc_kernel), DESIGN.md "Kernels":
# per sample and epoch: forward + backward-data + backward-weight of both 14-64-64 towers
FLOP_PER_SAMPLE_EPOCH = 61056  # SURVEY.md section 8d, point
ENV_STEP_BYTES = 145           # SURVEY.md section 8d, stand-alone point env-step


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-steps", type=int, default=148, help="rollout length of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.lines, self.index = None, [], index
        self.active = True   # only samples taken while a timed region runs are kept

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            if self.active:
                self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


# ------------------------------------------------------------------------------------------
def cpu_port_step(n_envs, n_steps, batch, n_epochs, state):
    """One iteration of the reference algorithm on the host: oracle env + torch-CPU PPO."""
    import numpy as np
    import torch

    from oracle import sb3_oracle

    ro, pol, opt = state["ro"], state["pol"], state["opt"]
    eps = torch.randn((n_steps, n_envs, 2), generator=state["gen"]).numpy()
    buf = ro.collect(eps)
    sb3_oracle.train_epochs(pol, opt, buf, n_epochs, batch, rng=state["rng"], clip_range=0.2, ent_coef=0.05,
                            vf_coef=0.5)
    return n_envs * n_steps


def cpu_port_setup(n_envs, n_steps, seed=0):
    import numpy as np
    import torch

    from oracle import point_oracle as po, sb3_oracle
    from oracle.vec_oracle import GoalVecOracle

    torch.manual_seed(seed)
    pol = sb3_oracle.MlpPolicyOracle(14)
    venv = GoalVecOracle(po.PointBody(n_envs), seed=seed, time_limit=1000, terminate_on_goal=True)
    ro = sb3_oracle.RolloutOracle(venv, pol, n_steps, gamma=0.99, gae_lambda=0.5)
    return dict(ro=ro, pol=pol, opt=sb3_oracle.make_adam(pol), gen=torch.Generator().manual_seed(seed),
                rng=np.random.RandomState(seed))


def run_reference(args):
    """CPU arm: bounded sample of the same workload per step (same envs, same minibatch size and
    epochs, shorter rollout) with every host thread torch / numpy will use."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    T = args.cpu_sample_steps
    batch = min(BATCH, N_ENVS * T)
    state = cpu_port_setup(N_ENVS, T)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_port_step(N_ENVS, T, batch, N_EPOCHS, state)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        done += cpu_port_step(N_ENVS, T, batch, N_EPOCHS, state)
    dt = time.perf_counter() - t0
    value = done / dt
    sample = (f"{N_ENVS} envs x {T} steps per step ({N_ENVS * T} env-steps), {N_EPOCHS} epochs of "
              f"minibatch {batch}; numpy fp64 oracle env + torch-CPU PPO (SB3 arithmetic)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 physics / f32 PPO",
            "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from mobrob_b200 import _lib
    from mobrob_b200.rl_control.ppo import PPOCtrl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mobrob_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    cfg = dict(env_name="point", time_limit=1000, n_envs=N_ENVS, vec_env_type="dummy", enable_gui=False, seed=0,
               ppo_kwargs=dict(policy="MlpPolicy", n_steps=N_STEPS, n_epochs=N_EPOCHS, ent_coef=0.05,
                               gae_lambda=0.5, batch_size=BATCH, verbose=0, host_permutation=False))
    ctrl = PPOCtrl.from_config(cfg)
    ctrl.ppo.tensorboard_log = None
    model = ctrl.ppo
    steps_per_iter = N_ENVS * N_STEPS  # per rank

    def iteration():
        model.collect_rollouts()
        model.train()

    for _ in range(max(args.warmup, 3)):
        iteration()
    barrier()

    # ---- value: device-resident, CUDA events, max over ranks -----------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        iteration()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = _lib.launch_count() - launches0
    value = steps_per_iter * world * args.steps / (ms / 1e3)
    sampler.active = False

    # ---- phase split and roofline of the dominant kernel (live CUDA events, same stream) -----------
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    barrier()
    ev[0].record()
    model.collect_rollouts()
    ev[1].record()
    model.train()
    ev[2].record()
    barrier()
    rollout_ms, train_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    epoch_ms = time_epoch_kernel(model, dev)
    peaks, peaks_src = measured_peaks()
    fp32_peak = 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    tensor_peak = peaks.get("bf16_tflops", 1590.0)   # burst figure: the kernel is timed alone
    achieved = FLOP_PER_SAMPLE_EPOCH * steps_per_iter / (epoch_ms * 1e-3) / 1e12
    roofline = {"kernel": "ppo_epoch_tc_kernel<16> (one launch = one epoch = 64 minibatch updates: tcgen05 forward/"
                          "backward GEMMs, gradient reduction, clip, Adam)",
                "bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": achieved / tensor_peak, "traffic": EPOCH_KERNEL_DRAM_BYTES, "ms_per_launch": epoch_ms,
                "algorithmic": f"{FLOP_PER_SAMPLE_EPOCH} FLOP/sample/epoch (SURVEY 8d) x {steps_per_iter} samples per launch",
                "peak_source": f"bf16_tflops of MEASURED_PEAKS.json ({peaks_src})",
                "tensor_flops_issued_per_algorithmic_flop": 3,
                "frac_issued": 3 * achieved / tensor_peak,
                "frac_of_fp32_fma_peak": achieved / fp32_peak,
                "note": "fp32 products are formed from two fp16 parts per operand (3 MMAs per product, gradients within "
                        "1e-6 of torch fp32), so 3x the algorithmic FLOPs go through the tensor pipe; the kernel is bound by the "
                        "per-minibatch dependency chain (3 grid barriers + L2 hand-offs), see profiles/",
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, profiles/ (per launch)",
                "step_share": {"rollout_ms": rollout_ms, "update_ms": train_ms,
                               "epoch_kernel_ms_x_epochs": epoch_ms * N_EPOCHS}}
    env_roof = time_env_step_kernel(dev, peaks)

    # ---- e2e: the public API (PPOCtrl.learn) with SB3's host-side permutations (pinned H2D) and
    #      per-iteration D2H of the training statistics / episode buffer ---------------------------------
    model.permutation = "pool"   # host-drawn permutations (permfeed.py), pinned -> device every epoch
    model.verbose = 0
    barrier()
    model.learn(total_timesteps=steps_per_iter * world * 3, reset_num_timesteps=True)  # warm (logger paths too)
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    model.learn(total_timesteps=steps_per_iter * world * args.steps, reset_num_timesteps=True)
    barrier()
    wall = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    e2e_value = steps_per_iter * world * args.steps / float(wall.item())
    clocks = sampler.stop() if rank == 0 else None   # sampled every 20 ms inside the value and e2e timed regions
    h2d = N_EPOCHS * steps_per_iter * 8  # int64 permutations, pinned -> device, per rank
    d2h = EP_D2H_BYTES + 11 * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {"metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 physics / f32 policy+PPO", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_envs_per_gpu": N_ENVS, "n_steps": N_STEPS, "batch_size": BATCH,
                       "n_epochs": N_EPOCHS, "minibatches_per_epoch": steps_per_iter // BATCH,
                       "parallelism": f"dp{world} (envs sharded, gradient all-reduce)",
                       "l2": "no flush: every step rewrites its 102 MB rollout working set and reads 97 MB of "
                             "fresh permutations (> 126 MB L2 together)"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world,
                    "what": "PPOCtrl.learn(): minibatch permutations drawn on the host (thread pool, one "
                            "iteration ahead) and copied from pinned memory every epoch; logger + episode-buffer "
                            "reads (D2H) every iteration"},
            "roofline": roofline, "roofline_env_step": env_roof}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


EP_D2H_BYTES = 100 * 12 + 8


# dram bytes of one ppo_epoch_tc_kernel launch (ncu --set full, see profiles/README.md)
EPOCH_KERNEL_DRAM_BYTES = 149.8e6  # 145.5 MB read + 4.2 MB written (profiles/r01_ncu_full_final.txt)
ENV_STEP_DRAM_BYTES = 808.1e6      # 318.8 MB read + 489.3 MB written at 2^22 envs (profiles/r01_env_step_ncu_final.txt)


def time_epoch_kernel(model, dev):
    """Average duration of the epoch kernel alone, CUDA events on the launching stream.  Parameters
    and Adam state are restored afterwards (the launches are real updates)."""
    import torch

    up, b = model.updater, model.buf
    T, N = model.n_steps, model.env.num_envs
    saved = [t.clone() for t in (up.params, up.exp_avg, up.exp_avg_sq, up.step)]
    perm = torch.randperm(T * N, device=dev, dtype=torch.int64)
    stats = up.adv_stats(b["advantages"], perm, BATCH, N, T)
    for _ in range(2):
        up.train_epoch_fused(b, perm, stats, BATCH, N, T)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        up.train_epoch_fused(b, perm, stats, BATCH, N, T)
    e1.record()
    torch.cuda.synchronize(dev)
    for t, sv in zip((up.params, up.exp_avg, up.exp_avg_sq, up.step), saved):
        t.copy_(sv)
    return e0.elapsed_time(e1) / reps


def time_env_step_kernel(dev, peaks):
    """Stand-alone env-step kernel at N = 2^22 envs (state >> L2): HBM roofline of K1."""
    import torch

    from mobrob_b200 import GpuVecEnv

    n = 1 << 22
    env = GpuVecEnv("point", n, seed=None, time_limit=1000, terminate_on_goal=True, device=dev.index)
    import numpy as np

    # cheap synthetic seeding for 4M envs: replicate a block of real PCG64 streams
    from mobrob_b200 import seeding, _lib as L

    blk = 4096
    init, goal, eng = seeding.vec_env_streams(0, blk)
    reps = n // blk
    init = np.tile(init, (reps, 1)); goal = np.tile(goal, (reps, 1))
    init[:, 1] += np.repeat(np.arange(reps, dtype=np.uint64), blk) * np.uint64(2654435761)
    goal[:, 1] += np.repeat(np.arange(reps, dtype=np.uint64), blk) * np.uint64(40503)
    eng = np.arange(n, dtype=np.int64)
    L.check(env.lib.mr_env_seed(env._h, L.ptr(init), L.ptr(goal), L.ptr(eng), env._stream()))
    env.reset_tensor()
    act = (torch.rand((n, 2), device=dev) * 2 - 1).sign().contiguous()
    for _ in range(3):
        env.step_tensor(act)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        env.step_tensor(act)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    achieved = ENV_STEP_BYTES * n / (ms * 1e-3) / 1e9
    env.close()
    return {"kernel": "point_step_kernel", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
            "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": ENV_STEP_DRAM_BYTES, "ms_per_launch": ms,
            "n_envs": n, "env_steps_per_s": n / (ms * 1e-3),
            "note": "algorithmic 145 B/env-step (SURVEY 8d); the state is fp64, physical traffic is 198 B/env-step (76 read, 122 written)"}


def cpu_baseline(args):
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    T = args.cpu_sample_steps
    batch = min(BATCH, N_ENVS * T)
    state = cpu_port_setup(N_ENVS, T)
    t0 = time.perf_counter()
    done = cpu_port_step(N_ENVS, T, batch, N_EPOCHS, state)
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"one iteration of {N_ENVS} envs x {T} steps ({done} env-steps), {N_EPOCHS} epochs of "
                      f"minibatch {batch}; numpy fp64 oracle env + torch-CPU PPO; {dt:.1f} s",
            "reference_stack_historical": "1026 env-steps/s (SB3 + mujoco-py, 2 subproc envs; BASELINE.md)"}


def _claim_stdout():
    """The contract is ONE JSON line on stdout: keep a private handle to the real stdout and point
    fd 1 at stderr, so that anything a library prints (NCCL's version banner, torch warnings)
    cannot land in front of it."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


if __name__ == "__main__":
    _claim_stdout()
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
