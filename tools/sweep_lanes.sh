#!/usr/bin/env bash
# Heavy-batch width against the number of envs (GPU box).  usage: sweep_lanes.sh "<sizes>" "<lanes>"
set -u
SIZES=${1:-"65536 262144"}
LANES=${2:-"32 16 8 4 2 1"}
for n in $SIZES; do for l in $LANES; do
  echo -n "envs $n lanes $l: "
  MSOC_HEAVY_LANES=$l python bench.py --steps 300 --warmup 20 --preroll 1000 --e2e-steps 1 --no-cpu-baseline --no-extras --envs-per-gpu $n 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4f ms' % d['ms_per_step'])"
done; done | tee gpurun_out/sweep_lanes_$(date +%H%M%S).log
