#!/bin/bash
# skewed two-tile pipeline in k_app_fwd_tc: parity, per-call times, ncu source-level captures of the dynamic tile kernels
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_paths.py -m gpu -q --timeout 120 -x > $O/r02f_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02f_pytest.log
tail -4 $O/r02f_pytest.log
timeout 200 python tools/exp_calls.py > $O/r02f_calls_new.txt 2> $O/r02f_calls_new.err
cat $O/r02f_calls_new.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:^k_app_fwd_tc --launch-skip 2 -c 2 -f -o $O/r02f_k_app_fwd_tc python tools/prof_bwd.py > $O/r02f_ncu1.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:^k_dyn_sigma_blend_tc2 --launch-skip 1 -c 1 -f -o $O/r02f_k_dyn_sigma_blend_tc2 python tools/prof_bwd.py > $O/r02f_ncu2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:^k_dyn_mlp_bwd_tc --launch-skip 1 -c 1 -f -o $O/r02f_k_dyn_mlp_bwd_tc python tools/prof_bwd.py > $O/r02f_ncu3.log 2>&1
ls -la $O/r02f*
