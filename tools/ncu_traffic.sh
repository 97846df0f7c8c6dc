#!/usr/bin/env bash
# DRAM / L2 traffic of the step kernel per variant and workload (GPU box).  usage: ncu_traffic.sh <which> <lib>...
set -u
W=$1; shift
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum
for lib in "$@"; do
  if [ "$lib" = base ]; then unset MSOC_LIB; else export MSOC_LIB=$PWD/$lib; fi
  echo "== $lib / $W"
  ncu --metrics $M --clock-control none -k regex:msoc_step -s 14 -c 1 python tools/exp.py --which $W --steps 6 2>&1 | grep -E "dram__|gpu__time|lts__|l1tex__" 
done
