#!/usr/bin/env bash
# Round records beside tools/gpu_measure.sh (GPU box): step time against the number of envs, the timeline of the contact
# kernel's batches by class (needs scratch/ab/libmsoc_tl.so = the library built with -DMSOC_TIMELINE), HBM reference points.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
{
  for n in 16384 65536 131072 262144 524288 1048576; do
    python bench.py --steps 300 --warmup 20 --preroll 1000 --e2e-steps 1 --no-cpu-baseline --no-extras --envs-per-gpu $n 2>&1 | tail -1 |
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('envs %8d  %.4f ms per step  %.4g env-steps/s' % (d['config']['envs_per_gpu'], d['ms_per_step'], d['value']))"
  done
} | tee gpurun_out/${TAG}_sizes.txt
if [ -f scratch/ab/libmsoc_tl.so ]; then
  for n in 1048576 65536; do
    echo "== timeline of one step, $n envs (us from the start of the streaming kernel; one record per warp-batch)"
    MSOC_LIB=$PWD/scratch/ab/libmsoc_tl.so python tools/timeline.py $n
  done | tee gpurun_out/${TAG}_timeline.txt
fi
python tools/hbm_mix.py | tee gpurun_out/${TAG}_hbm_mix.json
