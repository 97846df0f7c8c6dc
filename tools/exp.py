#!/usr/bin/env python
"""Kernel limit experiments (GPU box): times the step kernel on workloads that isolate parts of it.
    python tools/exp.py [--envs N] [--steps K]
  still    default random spawn, zero actions: no contact ever -> pure contact-free path + observation writer
  bench    the bench.py mix (short pre-roll) for reference
"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from marl_soccer_b200 import _capi
from marl_soccer_b200.sim import BatchedSoccerSim, load_default_config

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--which", default="still,open")
args = ap.parse_args()
dev = torch.device("cuda:0")
n = args.envs
cfg = load_default_config()
sim = BatchedSoccerSim(n, config=cfg, device=dev, seed=0)


def timed(name, act_fn, steps):
    for k in range(10):
        sim.step(act_fn(k))
    sim.stats(reset=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        sim.step(act_fn(k))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = sim.stats()
    print(f"{name}: {ms:.4f} ms/step  {n / ms * 1e-6:.3f}e9 env-steps/s  frac {n * 2194 / (ms * 1e-3) / 6546.6e9:.3f}  "
          f"contacts/env-step {st['contacts'] / max(st['env_steps'], 1):.3f}", flush=True)


zero = torch.zeros((n, 4, 3), device=dev)
small = (torch.rand((4, n, 4, 3), device=dev) * 2 - 1)
for w in args.which.split(","):
    if w == "still":
        sim.reset(_capi.MODE_RANDOM, seed=1)
        timed("still (no contacts, zero actions)", lambda k: zero, args.steps)
    elif w == "open":
        sim.reset(_capi.MODE_RANDOM, seed=2)
        timed("open (random spawn, random actions, first 200 steps)", lambda k: small[k % 4], args.steps)
    elif w == "full":
        sim.reset(_capi.MODE_FULL_RANDOM, seed=3)
        timed("full-random spawn, random actions, first steps", lambda k: small[k % 4], args.steps)
