#!/bin/bash
# A/B on ONE box: tools/ab.sh <libA> <libB> [rounds] -- alternates two builds of libaninerf_b200.so: stage timers (ms, clock-dependent)
# and the traced build's cycles per tile (clock-independent)
A=$1; B=$2; R=${3:-2}
cp animatable_nerf_b200/libaninerf_b200.so /tmp/lib_keep.so
for r in $(seq $R); do
  for v in $A $B; do
    cp $v animatable_nerf_b200/libaninerf_b200.so
    echo "== $v"; timeout 300 python tools/gpu_diag.py mlp_time 2>&1 | grep "mlp_time:"
    ANINERF_TRACE_ITER=5 timeout 300 python tools/gpu_diag.py trace_frame 2>&1 | grep "mean period" | cut -c1-72
  done
done
cp /tmp/lib_keep.so animatable_nerf_b200/libaninerf_b200.so
