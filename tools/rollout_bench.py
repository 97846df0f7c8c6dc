#!/usr/bin/env python
"""BASELINE config 5 on ONE GPU's shard: PPO self-play rollout loop, 262 144 envs x 128 steps with the reference's
torch MLP policy (66->512->256->128->64->3, tanh) in the loop (marl_soccer_b200/rollout.py).  Prints env-steps/s of
the whole rollout (simulator + normaliser + policy + storage) and of the simulator alone on the same states.
    python tools/rollout_bench.py [--envs N] [--steps T] [--dtype bf16|fp32]
Under torchrun every rank runs its shard (envs sharded by global index, no per-step collective) and rank 0
prints the aggregate after one all-reduce of the rollout statistics."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from marl_soccer_b200 import _capi
from marl_soccer_b200.sim import BatchedSoccerSim, load_default_config
from marl_soccer_b200.rollout import Agent, RolloutBuffer, RunningMeanStd, collect_rollout

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=262144)
ap.add_argument("--steps", type=int, default=128)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--rollouts", type=int, default=3)
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
n, T = args.envs, args.steps
sim = BatchedSoccerSim(n, config=load_default_config(), device=dev, seed=0, global_env_offset=rank * n)
torch.manual_seed(0)
agent = Agent().to(dev)
rms = RunningMeanStd((66,), dev)
obs_dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
buf = RolloutBuffer(T, n, dev, obs_dtype=obs_dtype)  # (T, N, 2, 66): 8.9 GB in fp32 at the default size
obs = sim.reset(_capi.MODE_FULL_RANDOM, seed=0)[:, :2].clone()
done = torch.zeros((n, 2), device=dev)
pdt = torch.bfloat16 if args.dtype == "bf16" else None
gen = torch.Generator(device=dev).manual_seed(1 + rank)
obs, done = collect_rollout(sim, agent, rms, buf, obs, done, generator=gen, policy_dtype=pdt)  # warm-up
sim.stats(reset=True)
torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.rollouts):
    obs, done = collect_rollout(sim, agent, rms, buf, obs, done, generator=gen, policy_dtype=pdt)
stats = sim.stats_tensor()
if dist is not None:
    dist.all_reduce(stats)  # the only collective: episode return / goal statistics of the rollout
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
t = torch.tensor([ms], device=dev, dtype=torch.float64)
if dist is not None:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
# simulator alone on the same kind of states
# i.i.d. actions: random windows of one flat buffer (a short cycle of tensors gives every env a periodic sequence
# with a net drift that pins the agents against the walls)
import numpy as np
flat = torch.rand((16 * n * 12,), device=dev) * 2 - 1
offs = np.random.default_rng(5).integers(0, 15 * n, size=T) * 12
for k in range(T):  # the rollout above ended in its own state mix; settle into the random-action mix first
    sim.step(flat[int(offs[k]):int(offs[k]) + n * 12].view(n, 4, 3))
torch.cuda.synchronize()
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
for k in range(T):
    sim.step(flat[int(offs[T - 1 - k]):int(offs[T - 1 - k]) + n * 12].view(n, 4, 3))
s1.record(); torch.cuda.synchronize()
if rank == 0:
    st = dict(zip([k for k, _ in _capi.MsocStats._fields_], stats.cpu().tolist()))
    print(json.dumps({"workload": f"PPO rollout, {n} envs x {T} steps per GPU, MLP policy in the loop ({args.dtype})", "n_gpus": world,
                      "rollout_env_steps_per_s": world * n * T * args.rollouts / (ms * 1e-3), "ms_per_rollout_step": ms / (T * args.rollouts),
                      "sim_only_env_steps_per_s_per_gpu": n * T / (s0.elapsed_time(s1) * 1e-3), "stats": st}))
if dist is not None:
    dist.destroy_process_group()
