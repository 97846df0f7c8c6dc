#!/usr/bin/env bash
# Kernel A/B on a B200 box (through gpurun): short bench of every library variant given on the command
# line (paths relative to the repo root; "base" = the in-tree libmsoc.so).  One line per variant.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/ab_$(date +%H%M%S).log
ARGS=${AB_ARGS:---steps 300 --warmup 20 --preroll 1000 --e2e-steps 1 --no-cpu-baseline}
for lib in "$@"; do
  if [ "$lib" = base ]; then unset MSOC_LIB; else export MSOC_LIB=$PWD/$lib; fi
  r=$(python bench.py $ARGS 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4g env-steps/s  %.4f ms  frac %.3f  contacts/step %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['stats']['contacts']/d['stats']['env_steps']))" 2>&1)
  echo "$lib: $r" | tee -a $OUT
done
