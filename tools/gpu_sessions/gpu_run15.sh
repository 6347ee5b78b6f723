#!/bin/bash
# two-tile sigma_blend + paired gathers + hoisted encodings + index prefetch: suite, per-call times, ncu of the backward tile kernels
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x > $O/r02h_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02h_pytest.log
tail -4 $O/r02h_pytest.log
timeout 200 python tools/exp_calls.py > $O/r02h_calls_new.txt 2> $O/r02h_calls_new.err
cut -c1-600 $O/r02h_calls_new.txt
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02h_train_new.json 2> $O/r02h_train_new.err
timeout 200 python bench.py --mode fwd --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02h_fwd_new.json 2> $O/r02h_fwd_new.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:^k_head_bwd_tc --launch-skip 2 -c 2 -f -o $O/r02h_k_head_bwd_tc python tools/prof_bwd.py > $O/r02h_ncu1.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:^k_dyn_mlp_bwd_tc --launch-skip 1 -c 1 -f -o $O/r02h_k_dyn_mlp_bwd_tc python tools/prof_bwd.py > $O/r02h_ncu2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:^k_dyn_sigma_blend_tc2 --launch-skip 1 -c 1 -f -o $O/r02h_k_dyn_sigma_blend_tc2 python tools/prof_bwd.py > $O/r02h_ncu3.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02h_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), d.get('kernel_ms_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY
