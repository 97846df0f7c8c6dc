#!/usr/bin/env python
"""Behavioural cross-check against the TRUE pymunk dynamics (SURVEY.md section 8c, weak check 2).

The reference ships policies trained on the real pymunk-backed env (runs/runN/*.ppo_model + the observation
normaliser).  A policy encodes the dynamics it was trained on: driven exactly as eval.py:84-104 does (blue =
deterministic policy on normalised observations, red = uniform random), it should behave on a faithful
simulator the way it did in training -- e.g. run4's final training average of ~120 blue return per episode
(charts/avg_agent0_return) means it scores readily.  This script plays episodes on the CPU
oracle (or, with --backend device, on the CUDA kernels) and prints returns, scores and the observation
statistics next to the shipped running mean.  Needs /root/reference (it reads the checkpoints), so it is a
tool, not a test; results are recorded in DESIGN.md."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/soccer_simulation"


def build_actor():
    return nn.Sequential(nn.Linear(66, 512), nn.Tanh(), nn.Linear(512, 256), nn.Tanh(), nn.Linear(256, 128), nn.Tanh(),
                         nn.Linear(128, 64), nn.Tanh(), nn.Linear(64, 3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--run", default="run5")
    ap.add_argument("--episodes", type=int, default=64)
    ap.add_argument("--backend", default="oracle", choices=["oracle", "hostsim", "device"])
    ap.add_argument("--mode", type=int, default=0, help="0 default random spawn (eval.py), 2 full random (training auto-reset)")
    ap.add_argument("--eval-weights", action="store_true", help="use the *.eval files of the run")
    args = ap.parse_args()
    import oracle_lib as O
    import parity_util as P
    sfx = ".eval" if args.eval_weights else ""
    sd = torch.load(os.path.join(REF, "runs", args.run, "ppo_pettingzoo_soccer.ppo_model" + sfx), map_location="cpu")
    actor = build_actor()
    actor.load_state_dict({k[len("actor_mean."):]: v for k, v in sd.items() if k.startswith("actor_mean.")})
    actor.eval()
    nz = np.load(os.path.join(REF, "runs", args.run, "latest_normalizer_stats_eval.npz" if args.eval_weights else "latest_normalizer_stats.npz"))
    mean, std = nz["mean"].astype(np.float64), np.sqrt(nz["var"].astype(np.float64))
    n = args.episodes
    if args.backend == "oracle":
        sim = O.OracleVec(n, P.CONFIG, seed=1)
    elif args.backend == "hostsim":
        import hostsim_lib as H
        sim = H.HostSim(n, P.CONFIG, seed=1)
    else:
        from marl_soccer_b200.host_api import HostBufferSim
        sim = HostBufferSim(n, P.CONFIG, seed=1)
    obs = sim.reset(args.mode, seed=1)
    rng = np.random.default_rng(0)
    ret = np.zeros(n)
    goals = np.zeros((n, 2), int)
    acc, cnt = np.zeros(66), 0
    for t in range(1000):
        blue = obs[:, :2].reshape(-1, 66).astype(np.float64)
        acc += blue.sum(0); cnt += blue.shape[0]
        x = torch.tensor(np.clip((blue - mean) / (std + 1e-8), -10, 10), dtype=torch.float32)
        with torch.no_grad():
            a_blue = actor(x).numpy().reshape(n, 2, 3)
        act = np.concatenate([a_blue, rng.uniform(-1, 1, (n, 2, 3))], axis=1).astype(np.float32)
        out = sim.step(act, auto_reset=False)
        obs, rew, done, goal = out[0], out[1], out[2], out[3]
        ret += np.asarray(rew)[:, 0]
        goals[:, 0] += goal > 0
        goals[:, 1] += goal < 0
    print(f"{args.run}{sfx} on {args.backend}, spawn mode {args.mode}: {n} episodes x 1000 steps")
    print(f"  blue return per episode: mean {ret.mean():.2f}  median {np.median(ret):.2f}  min {ret.min():.2f}  max {ret.max():.2f}")
    print(f"  goals per episode: blue {goals[:, 0].mean():.2f}  red {goals[:, 1].mean():.2f}")
    m = acc / cnt
    idx = [0, 1, 2, 3, 13, 14, 15, 16, 18, 19, 21]
    print("  newest-frame obs mean (ours | shipped running mean) for features", idx)
    print("   ", np.round(m[44:][idx], 3))
    print("   ", np.round(mean[44:][idx], 3))


if __name__ == "__main__":
    main()
