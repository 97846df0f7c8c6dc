#!/usr/bin/env python
"""bench.py -- env-steps/sec of the fused soccer step on N B200s (weak scaling, envs sharded by
global env index, no per-step collective), with the HBM roofline of the step, the end-to-end number
through the host-buffer C-ABI call, and the CPU oracle timed on the host cores as the reported baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu E] [--impl reference]

One "step" = one msoc_step call over this rank's shard = E env-steps = three kernel launches (the streaming
contact-free kernel over all envs, then the light and the heavy contact kernels side by side on two
streams).  Under torchrun (N > 1) every rank owns E envs with global indices [rank*E, (rank+1)*E).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_ENV_STEP = 2194  # 940 B read + 1 254 B written, fp32, drop-in semantics
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
DEFAULT_ENVS_PER_GPU = 1_048_576  # BASELINE config 3 size; working set >> 126 MB L2
WORKLOAD = ("2v2 soccer, {n} envs per GPU (BASELINE config-3 size on every GPU), config.json defaults, "
            "i.i.d. random actions U(-1,1) (random windows of a pre-generated device buffer), auto-reset in full-random mode, shaped "
            "rewards, episodes staggered over all phases")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per step (sum over the three step kernels) from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def cpu_oracle_run(n_envs: int, steps: int, threads: int, seed: int = 0):
    """Times the CPU oracle (oracle/liboracle.so, the restated reference path) on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib as O
    vec = O.OracleVec(n_envs, O.DEFAULT_CONFIG, seed=seed)
    vec.reset(O.MODE_FULL_RANDOM, seed=seed)
    rng = np.random.default_rng(seed)
    flat = rng.uniform(-1, 1, 16 * n_envs * 12).astype(np.float32)  # i.i.d. windows, as on the GPU arm
    offs = rng.integers(0, 15 * n_envs, size=steps + 1) * 12
    vec.step(flat[offs[steps]:offs[steps] + n_envs * 12], auto_reset=True, nthreads=threads)
    t0 = time.perf_counter()
    for k in range(steps):
        vec.step(flat[offs[k]:offs[k] + n_envs * 12], auto_reset=True, nthreads=threads)
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  pymunk/pygame/gymnasium/
    pettingzoo are not installable here (no network) and the reference is pure Python over them, so the
    arm times the oracle port (oracle/, the CPU restatement) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_envs, per_step = 1024, 1000  # one bench "step" = one full 1000-step episode of 1024 envs (all phases)
    vals = []
    for _ in range(args.warmup):
        cpu_oracle_run(n_envs, 50, cores)
    t_total = 0.0
    for _ in range(args.steps):
        v, dt = cpu_oracle_run(n_envs, per_step, cores)
        vals.append(v)
        t_total += dt
    value = n_envs * per_step * len(vals) / t_total
    sample = f"{n_envs} envs x {per_step} steps per bench step, {len(vals)} bench steps, OpenMP over {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, len(vals)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(n=args.envs_per_gpu)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--envs-per-gpu", type=int, default=DEFAULT_ENVS_PER_GPU)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preroll", type=int, default=1000, help="untimed steps that decorrelate the episode phases")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 20:
            args.steps = 5
        args.warmup = min(args.warmup, 1)
        run_reference(args)
        return

    import numpy as np
    import torch
    from marl_soccer_b200 import _capi
    from marl_soccer_b200.sim import BatchedSoccerSim, load_default_config

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs CUDA devices (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version banner there)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs_per_gpu
    cfg = load_default_config()
    L = _capi.lib()
    sim = BatchedSoccerSim(n, config=cfg, device=dev, seed=0, global_env_offset=rank * n)
    sim.reset(_capi.MODE_FULL_RANDOM, seed=0)
    # Actions: one flat device buffer of 16*N*12 uniform(-1,1) floats generated before the timed region;
    # step k reads the window starting at a pseudo-random env offset r_k, so every env sees an
    # effectively i.i.d. action stream (a plain cycle over 16 tensors would give each env a periodic
    # sequence with a constant net drift, pinning the agents against the walls).
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    POOL = 16
    flat = torch.rand((POOL * n * 12,), generator=gen, device=dev) * 2 - 1
    offs = np.random.default_rng(99 + rank).integers(0, (POOL - 1) * n, size=1 << 16)

    class _Pool:
        def __getitem__(self, k):
            o = int(offs[k % len(offs)]) * 12
            return flat[o:o + n * 12].view(n, 4, 3)
    pool = _Pool()

    # pre-roll (untimed): env i is re-spawned at pre-roll step hash(i) % max_steps, so after one episode
    # length the episode phases are uniformly staggered and any timed window sees the time-average mix of
    # spawn overlap, resting contacts, goals and auto-resets, whatever --steps is.
    max_steps = int(cfg["simulation"]["max_steps"])
    phase = (torch.arange(n, device=dev, dtype=torch.int64) * 2654435761) % max_steps
    for k in range(args.preroll):
        sim.reset(_capi.MODE_FULL_RANDOM, mask=(phase == (k % max_steps)))
        sim.step(pool[k])
    for k in range(max(3, args.warmup)):
        sim.step(pool[args.preroll + k])
    sim.stats(reset=True)
    torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches0 = L.msoc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    base_k = args.preroll + max(3, args.warmup)
    for k in range(args.steps):
        sim.step(pool[base_k + k])
    stats_t = sim.stats_tensor(reset=False)
    if dist is not None:
        dist.all_reduce(stats_t)  # the only collective: one 64-byte sum per rollout
    e1.record()
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = L.msoc_launch_count() - launches0
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    stats = dict(zip([k for k, _ in _capi.MsocStats._fields_], stats_t.cpu().tolist()))

    # device time of one step on this rank (three launches; events bracket K back-to-back steps)
    kernel_ms = e0.elapsed_time(e1) / args.steps
    value = world * n * args.steps / (elapsed_ms * 1e-3)

    # end to end through the host-buffer C-ABI call (msoc_step_host): pinned host actions in, pinned host
    # obs / reward / done / goal / score out, copies inside the timed region
    import ctypes as C
    h_act = torch.empty((n, 4, 3), dtype=torch.float32).pin_memory()
    h_act.copy_(pool[0].cpu())
    h_acts = [h_act] + [torch.empty((n, 4, 3), dtype=torch.float32).pin_memory().copy_(pool[7 + j].cpu()) for j in range(3)]
    h_obs = torch.empty((n, 4, 66), dtype=torch.float32).pin_memory()
    h_rew = torch.empty((n, 2), dtype=torch.float32).pin_memory()
    h_done = torch.empty((n,), dtype=torch.uint8).pin_memory()
    h_goal = torch.empty((n,), dtype=torch.int8).pin_memory()
    h_score = torch.empty((n, 2), dtype=torch.int32).pin_memory()
    stream = torch.cuda.current_stream(dev).cuda_stream

    def host_step(j=0):
        _capi.check(L.msoc_step_host(sim._h, h_acts[j % 4].data_ptr(), h_obs.data_ptr(), h_rew.data_ptr(), h_done.data_ptr(),
                                     h_goal.data_ptr(), h_score.data_ptr(), _capi.STEP_AUTO_RESET, stream))
    for _ in range(3):
        host_step()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for j in range(args.e2e_steps):
        host_step(j)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.e2e_steps / float(t.item())
    h2d = n * 12 * 4
    d2h = n * (4 * 66 * 4 + 2 * 4 + 1 + 1 + 2 * 4)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    achieved = n * BYTES_PER_ENV_STEP / (kernel_ms * 1e-3) / 1e9
    traffic = recorded_traffic()
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        steps_cpu = 1000  # one full episode: spawn overlap, contacts, truncation + auto-reset (~10-20 s)
        v, dt = cpu_oracle_run(4096, steps_cpu, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"oracle/liboracle.so (restated reference, pymunk unavailable): 4096 envs x {steps_cpu} steps, "
                         f"OpenMP over {cores} threads, {dt:.1f} s"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(n=n), "envs_per_gpu": n, "global_envs": world * n,
                   "l2": "working set per GPU (state + obs + actions, ~1.5 KB/env) >> 126 MB L2; no flush needed",
                   "preroll_steps": args.preroll},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (traffic or {}).get("dram_bytes_per_launch_at_bench_size"),
                     "peak_source": peak_src, "algorithmic_bytes_per_env_step": BYTES_PER_ENV_STEP,
                     "kernel": "msoc_step = msoc_step_fast_kernel + msoc_step_light_kernel || msoc_step_contact_kernel "
                               "(whole step: algorithmic bytes of all envs / device time of the three launches)",
                     "kernel_ms": kernel_ms},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": args.e2e_steps, "api": "msoc_step_host (pinned host buffers)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "stats": stats,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
