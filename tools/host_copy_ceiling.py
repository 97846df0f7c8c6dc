#!/usr/bin/env python
"""Host-side ceiling of the end-to-end path (GPU box, 1..8 ranks under torchrun): every rank moves exactly the bytes one
msoc_step_host_frames call moves (48 B/env of actions host->device, 370 B/env of frames, rewards and flags device->host)
between pinned host buffers and its GPU, with NO kernels in between, the same chunking on two streams.  The aggregate
env-steps/s it prints is what the copies alone allow on this host (PCIe root complexes, host memory, NUMA); the end-to-end
figure of bench.py can only approach it.

    python -m torch.distributed.run --nproc-per-node 8 ... tools/host_copy_ceiling.py [--envs-per-gpu N] [--steps K]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--no-bind", action="store_true")
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist = None
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
near = None
if not args.no_bind:
    _, near = bench.bind_near_gpu(torch, dev)
n = args.envs_per_gpu
h2d_b, d2h_b = n * 48, n * (4 * 22 * 4 + 8 + 1 + 1 + 8)
h_in = torch.empty(h2d_b, dtype=torch.uint8).pin_memory()
h_out = torch.empty(d2h_b, dtype=torch.uint8).pin_memory()
d_in = torch.empty(h2d_b, dtype=torch.uint8, device=dev)
d_out = torch.empty(d2h_b, dtype=torch.uint8, device=dev)
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
CH = 8


def step():
    for c in range(CH):
        with torch.cuda.stream(streams[c & 1]):
            a, b = c * h2d_b // CH, (c + 1) * h2d_b // CH
            d_in[a:b].copy_(h_in[a:b], non_blocking=True)
            a, b = c * d2h_b // CH, (c + 1) * d2h_b // CH
            h_out[a:b].copy_(d_out[a:b], non_blocking=True)
    for s in streams:
        s.synchronize()


for _ in range(3):
    step()
if dist is not None:
    dist.barrier()
torch.cuda.synchronize(dev)
t0 = time.perf_counter()
for _ in range(args.steps):
    step()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
if dist is not None:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    dt = float(dt.item())
    print(json.dumps({"what": "host copy ceiling of msoc_step_host_frames (copies only)", "n_gpus": world, "envs_per_gpu": n,
                      "env_steps_per_s": world * n * args.steps / dt, "gb_per_s_all_gpus": world * (h2d_b + d2h_b) * args.steps / dt / 1e9,
                      "gb_per_s_per_gpu": (h2d_b + d2h_b) * args.steps / dt / 1e9, "host_cpus_near_gpu": near, "bound_to_numa_node": not args.no_bind}))
if dist is not None:
    dist.destroy_process_group()
