#!/usr/bin/env bash
# One ncu --set full capture (with source counters) of the step kernels of one bench step (GPU box).  usage: ncu_full.sh <tag> [lib]
set -u
TAG=$1
if [ $# -ge 2 ] && [ "$2" != base ]; then export MSOC_LIB=$PWD/$2; fi
SHORT="python bench.py --steps 40 --warmup 5 --e2e-steps 2 --no-cpu-baseline --no-extras"
$SHORT > gpurun_out/${TAG}_plain.json 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:msoc_step -s 2020 -c 2 -f -o gpurun_out/${TAG}_full $SHORT > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
