#!/bin/bash
# A/B on ONE box: tools/ab.sh <libA> <libB> [rounds] -- alternates the two builds of libaninerf_b200.so under tools/gpu_diag.py mlp_time
A=$1; B=$2; R=${3:-2}
cp animatable_nerf_b200/libaninerf_b200.so /tmp/lib_keep.so
for r in $(seq $R); do
  for v in $A $B; do
    cp $v animatable_nerf_b200/libaninerf_b200.so
    echo "== $v"; timeout 300 python tools/gpu_diag.py mlp_time 2>&1 | grep "mlp_time:"
  done
done
cp /tmp/lib_keep.so animatable_nerf_b200/libaninerf_b200.so
