#!/usr/bin/env bash
# Multi-GPU measurement on one 8-GPU box (gpurun --gpus 8): weak scaling at 8 GPUs (1 Mi envs per GPU) and strong
# scaling of BASELINE config 4 as written (1 048 576 envs split over 1/2/4/8 GPUs).  One JSON line per run.
# STRONG_N / WEAK_N / CEIL_N (lists of GPU counts) select the runs, so that the 2- and 4-GPU points can be taken on smaller boxes.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
run() { # N extra-args... -> one line
  local n=$1; shift
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@"
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@"
  fi
}
COMMON="--steps 200 --warmup 20 --e2e-steps 5 --no-cpu-baseline --no-extras"
for n in ${STRONG_N-1 2 4 8}; do
  run $n $COMMON --global-envs 1048576 2> gpurun_out/${TAG}_strong_${n}gpu.err | tail -1 > gpurun_out/${TAG}_bench_strong_${n}gpu.json
  python -c "import json,sys; d=json.load(open('gpurun_out/${TAG}_bench_strong_${n}gpu.json')); print('strong', d['n_gpus'], '%.4g env-steps/s' % d['value'], '%.4f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], 'envs/gpu', d['config']['envs_per_gpu'])"
done
for n in ${WEAK_N-2 8}; do
  run $n $COMMON 2> gpurun_out/${TAG}_weak_${n}gpu.err | tail -1 > gpurun_out/${TAG}_bench_${n}gpu.json
  python -c "import json,sys; d=json.load(open('gpurun_out/${TAG}_bench_${n}gpu.json')); print('weak', d['n_gpus'], '%.4g env-steps/s' % d['value'], '%.4f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], 'full-obs e2e %.4g' % d['e2e']['full_observation']['value'])"
done
# what the host <-> device copies alone allow (no kernels): the ceiling of the end-to-end figure
for n in ${CEIL_N-1 8}; do
  if [ "$n" = 1 ]; then python tools/host_copy_ceiling.py; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) tools/host_copy_ceiling.py; fi 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_copy_ceiling_${n}gpu.json
done
