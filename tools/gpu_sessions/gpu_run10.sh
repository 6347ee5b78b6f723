#!/bin/bash
# ncu --set full with source-level stall samples for the five largest kernels of the train step (one launch each)
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
for K in k_dyn_sigma_blend_tc2 k_dyn_mlp_bwd_tc k_head_bwd_tc k_app_fwd_tc k_basis_bwd; do
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:^$K --launch-skip 2 -c 1 -f -o $O/r02c_$K python tools/prof_bwd.py > $O/r02c_ncu_$K.log 2>&1
done
ls -la $O/*.ncu-rep
