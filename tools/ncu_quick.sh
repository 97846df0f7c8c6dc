#!/usr/bin/env bash
# Per-kernel time / DRAM bytes / instructions / issue rate of one bench step (GPU box).  usage: ncu_quick.sh <tag> [lib]
set -u
TAG=$1
if [ $# -ge 2 ] && [ "$2" != base ]; then export MSOC_LIB=$PWD/$2; fi
SHORT="python bench.py --steps 40 --warmup 5 --e2e-steps 2 --no-cpu-baseline --no-extras"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum
ncu --metrics $M --clock-control none -k regex:msoc_step -s 2020 -c 2 --csv --log-file gpurun_out/${TAG}_quick.csv $SHORT > gpurun_out/${TAG}_quick.log 2>&1
python - <<PY
import csv
rows=list(csv.DictReader(l for l in open("gpurun_out/${TAG}_quick.csv") if l.startswith('"')))
out={}
for r in rows: out.setdefault(r["Kernel Name"].split("(")[0],{})[r["Metric Name"].split(".")[0].replace("smsp__","").replace("l1tex__t_sectors_pipe_lsu_mem_","")]=r["Metric Value"]
for k,v in out.items(): print(k, v)
PY
