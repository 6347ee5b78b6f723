#!/bin/bash
# stall-sample captures of the seven tile kernels of the FINAL build (one launch each, one process)
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:^k_(warp_fwd_tc2|dyn_sigma_blend_tc2|app_fwd_tc|head_bwd_tc|dyn_mlp_bwd_tc)' --launch-skip 7 -c 7 -f -o $O/r02k_tiles python tools/prof_bwd.py > $O/r02k_ncu.log 2>&1
ls -la $O/r02k_tiles.ncu-rep
