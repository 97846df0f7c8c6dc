#!/usr/bin/env bash
# Round measurement on a B200 box (run through gpurun): tests, full bench, reference arm, ncu launch list
# and one ncu --set full capture of one whole step (its two kernels).  Outputs land in gpurun_out/.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json
kill $SMI
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>&1; tail -c 300 gpurun_out/${TAG}_bench_reference.json
# the bench workload itself (same 1000-step pre-roll, so the same steady-state env mix), fewer timed steps
SHORT="python bench.py --steps 40 --warmup 5 --e2e-steps 2 --no-cpu-baseline --no-extras"
$SHORT > gpurun_out/${TAG}_short_plain.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:msoc_step -s 2020 -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv $SHORT > gpurun_out/${TAG}_ncu_launches.log 2>&1
$SHORT > gpurun_out/${TAG}_short_plain2.json 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:msoc_step -s 2020 -c 2 -f -o gpurun_out/${TAG}_step_full $SHORT > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out | tail -15
