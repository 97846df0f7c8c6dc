"""World-size-2 gloo test of the multi-GPU plumbing on CPU: shards keyed by global env index reproduce the
single-process run, and the per-rollout statistics all-reduce sums the shards."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marl_soccer_b200.distributed import allreduce_stats, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, steps, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    import hostsim_lib as H
    import parity_util as P
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    sim = H.HostSim(hi - lo, P.CONFIG, seed=3, global_offset=lo)
    sim.reset(2, seed=17)
    st = [sim.get_state(i) for i in range(hi - lo)]
    for s in st:
        s.steps = 1000 - steps + 2  # everybody truncates (and auto-resets) inside the rollout
    for i, s in enumerate(st):
        sim.set_state(i, s)
    rng = np.random.default_rng(5)
    acts = rng.uniform(-1, 1, (steps, n_total, 4, 3)).astype(np.float32)
    rews = []
    for t in range(steps):
        o, r, d, g = sim.step(acts[t, lo:hi])
        rews.append(r)
    stats = torch.tensor(list(sim.stats().values()), dtype=torch.float64)
    allreduce_stats(stats)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), obs=o, rew=np.stack(rews), stats=stats.numpy())
    dist.destroy_process_group()


def test_shard_range_partitions():
    for n, w in ((10, 3), (1048576, 8), (7, 8), (4096, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_two_rank_shards_match_single_process(tmp_path):
    import hostsim_lib as H
    import parity_util as P
    n_total, steps, world = 48, 12, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, steps, str(tmp_path)), nprocs=world, join=True)
    # single-process reference run
    sim = H.HostSim(n_total, P.CONFIG, seed=3)
    sim.reset(2, seed=17)
    for i in range(n_total):
        s = sim.get_state(i)
        s.steps = 1000 - steps + 2
        sim.set_state(i, s)
    rng = np.random.default_rng(5)
    acts = rng.uniform(-1, 1, (steps, n_total, 4, 3)).astype(np.float32)
    rews = []
    for t in range(steps):
        o, r, d, g = sim.step(acts[t])
        rews.append(r)
    ref_stats = np.array(list(sim.stats().values()))
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert np.array_equal(np.concatenate([p["obs"] for p in parts]), o)
    assert np.array_equal(np.concatenate([p["rew"] for p in parts], axis=1), np.stack(rews))
    for p in parts:  # every rank holds the global sums after the all-reduce
        assert np.allclose(p["stats"], ref_stats, rtol=1e-12, atol=1e-9)
    assert ref_stats[0] == n_total and ref_stats[4] == n_total * steps
