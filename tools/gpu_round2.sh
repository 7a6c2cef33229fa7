#!/bin/bash
# One GPU-box visit of round 2.  usage: tools/gpu_round2.sh <tag> [steps...]   steps: test bench mma stage san ncu
tag=${1:-x}; shift
steps=${@:-test bench}
mkdir -p gpurun_out
for s in $steps; do
case $s in
test)
  timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_$tag.log;;
testall)
  timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_$tag.log;;
bench)
  timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err;;
benchq)
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/benchq_$tag.json 2> gpurun_out/benchq_$tag.err; echo "benchq rc=$?"; cut -c1-2500 gpurun_out/benchq_$tag.json; tail -5 gpurun_out/benchq_$tag.err;;
ref)
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ref_$tag.json 2> gpurun_out/ref_$tag.err; echo "ref rc=$?"; cut -c1-800 gpurun_out/ref_$tag.json;;
verify)
  timeout 120 tools/bench_mma 1 verify > gpurun_out/verify_$tag.log 2>&1; echo "verify rc=$?"; cat gpurun_out/verify_$tag.log;;
mma)
  timeout 120 tools/bench_mma 1 > gpurun_out/mma1_$tag.log 2>&1; echo "mma1 rc=$?"; cat gpurun_out/mma1_$tag.log
  timeout 120 tools/bench_mma 74 > gpurun_out/mma74_$tag.log 2>&1; echo "mma74 rc=$?"; grep -v "^#" gpurun_out/mma74_$tag.log | head -40;;
stage)
  timeout 300 python tools/stage_kernels.py > gpurun_out/stage_$tag.log 2>&1; echo "stage rc=$?"; tail -4 gpurun_out/stage_$tag.log
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sample_bw_kernel|lbs_kernel|composite_kernel' --launch-skip 4 -c 17 -o gpurun_out/prof_stage_$tag -f python tools/stage_kernels.py > gpurun_out/ncu_stage_$tag.log 2>&1; echo "ncu stage rc=$?";;
san)
  for tool in memcheck synccheck racecheck; do
    timeout 420 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_frame.py > gpurun_out/san_${tool}_$tag.log 2>&1; echo "sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_FRAME_DONE|Error|hazard" gpurun_out/san_${tool}_$tag.log | head -8
  done;;
trace)
  timeout 300 python tools/gpu_diag.py trace_frame > gpurun_out/trace_$tag.log 2>&1; echo "trace rc=$?"; grep -A40 "trace_frame" gpurun_out/trace_$tag.log | head -90;;
dbg)
  for d in 1 2 3; do ANINERF_DEBUG_MLP=$d timeout 300 python tools/gpu_diag.py trace_frame > gpurun_out/trace_dbg${d}_$tag.log 2>&1; echo "dbg $d rc=$?"; grep -A22 "trace_frame" gpurun_out/trace_dbg${d}_$tag.log | grep "mean period\|L[1234] slot0"; done;;
ncu)
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_launch_$tag.log 2>&1; echo "ncu list rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mlp_kernel' --launch-skip 12 -c 4 -o gpurun_out/prof_$tag -f python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?";;
esac
done
