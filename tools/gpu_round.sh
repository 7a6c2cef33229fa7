#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list + full capture of the MLP kernels.
# usage: tools/gpu_round.sh <tag> [skip_ncu]
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$tag.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json | cut -c1-1500
if [ -z "$2" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_$tag.log 2>&1; echo "ncu list rc=$?"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'mlp_kernel|composite_kernel|mask_kernel' --launch-skip 24 -c 4 -o gpurun_out/prof_$tag -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
fi
