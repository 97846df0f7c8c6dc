#!/usr/bin/env bash
# compute-sanitizer memcheck + racecheck + synccheck of the step kernels (GPU box): smoke() and the ragged-size /
# masked-reset test.  racecheck covers the warp-shared contact pool and the staged observation blocks (shared memory).
set -u
TAG=${1:-r02}
OUT=gpurun_out/${TAG}_sanitizer.log
: > $OUT
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool: python -c 'import __graft_entry__ as g; g.smoke()'" >> $OUT
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^$" | tail -12 >> $OUT
  echo "=== compute-sanitizer --tool $tool: pytest tests/test_gpu_parity.py::test_ragged_sizes_and_masked_reset" >> $OUT
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py::test_ragged_sizes_and_masked_reset -x -q 2>&1 | grep -v "^$" | tail -12 >> $OUT
done
tail -60 $OUT
