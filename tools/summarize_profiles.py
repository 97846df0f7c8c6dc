#!/usr/bin/env python
"""Turns the raw outputs of tools/gpu_measure.sh (gpurun_out/<tag>_*) into the committed summaries under
profiles/: bench lines, ncu launch list, the key metrics of the ncu --set full capture of one step (its two kernels)
and profiles/traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic)."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

for name in ("bench.json", "bench_reference.json", "launches.csv", "clocks.csv", "short_plain.json", "pytest_gpu.log", "smoke.log"):
    src = os.path.join(G, f"{tag}_{name}")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, f"{tag}_{name}"))

rep = os.path.join(G, f"{tag}_step_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]
launches = []
names = []
for r in rows[2:]:
    launches.append({k: (rows[1][h.index(k)], r[h.index(k)]) for k in KEYS if k in h})
    names.append(r[h.index("Kernel Name")] if "Kernel Name" in h else "?")
with open(os.path.join(P, f"{tag}_step_kernel_ncu_full.txt"), "w") as f:
    f.write(f"ncu --set full --clock-control none --import-source on -k regex:msoc_step (tools/gpu_measure.sh {tag}); "
            "same command exited 0 without ncu first.  Cold-cache, serialised replays: use shares, not absolutes.\n")
    short = json.loads(open(os.path.join(G, f"{tag}_short_plain.json")).read().strip().splitlines()[-1])
    f.write(f"workload: {short['config']['envs_per_gpu']} envs per launch ({short['config']['workload']})\n\n")
    for i, L in enumerate(launches):
        f.write(f"--- launch {i}: {names[i]}\n")
        for k, (u, v) in L.items():
            f.write(f"{k:90s} {v} {u}\n")
def to_bytes(unit, v):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m[unit]
# one step = the two captured launches (fast, contact): DRAM traffic of the step = their sum
rd = sum(to_bytes(*L["dram__bytes_read.sum"]) for L in launches)
wr = sum(to_bytes(*L["dram__bytes_write.sum"]) for L in launches)
n = short["config"]["envs_per_gpu"]
traffic = {"tag": tag, "kernel": "msoc_step (" + " + ".join(names) + ")", "envs_per_launch": n, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_launch_at_bench_size": rd + wr, "dram_bytes_per_env_step": (rd + wr) / n,
           "algorithmic_bytes_per_env_step": 2194, "source": f"profiles/{tag}_step_kernel_ncu_full.txt"}
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(json.dumps(traffic))
