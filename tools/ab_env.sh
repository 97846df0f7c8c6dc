#!/usr/bin/env bash
# Kernel A/B on a B200 box (through gpurun): short bench under every environment setting given on the command line
# ("base" or a comma-separated list of VAR=value, e.g. MSOC_STEP_CHUNKS=4,MSOC_HEAVY_LANES=16).  One line per variant.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/abenv_$(date +%H%M%S).log
ARGS=${AB_ARGS:---steps 300 --warmup 20 --preroll 1000 --e2e-steps 1 --no-cpu-baseline --no-extras}
for spec in "$@"; do
  (
    if [ "$spec" != base ]; then IFS=',' read -ra KV <<< "$spec"; for kv in "${KV[@]}"; do export "$kv"; done; fi
    r=$(python bench.py $ARGS 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4g env-steps/s  %.4f ms  frac %.3f  contacts/step %.4f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['stats']['contacts']/d['stats']['env_steps']))" 2>&1)
    echo "$spec: $r" | tee -a $OUT
  )
done
