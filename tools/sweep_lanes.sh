#!/usr/bin/env bash
# heavy-kernel batch width against the number of envs (GPU box)
set -u
for n in 16384 65536 131072 262144 524288 1048576; do
  for hl in 8 16 32; do
    r=$(MSOC_HEAVY_LANES=$hl python bench.py --steps 300 --warmup 20 --preroll 1000 --e2e-steps 1 --no-cpu-baseline --no-extras --envs-per-gpu $n 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4f ms' % d['ms_per_step'])")
    echo "envs $n lanes $hl: $r" | tee -a gpurun_out/sweep_lanes.log
  done
done
