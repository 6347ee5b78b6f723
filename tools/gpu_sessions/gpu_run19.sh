#!/bin/bash
# final build, pass 2: per-kernel ncu table of one whole train step (written into profiles/ on the box so that the bench lines of
# this same call read it), every 1-GPU bench line, the tests added last, A/B of k_basis_bwd compiled for 3 CTAs per SM
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
KR='regex:^k_(dyn_mlp_bwd_tc|head_bwd_tc|basis_bwd|warp_bwd|app_fwd_tc|dyn_sigma_blend_tc2|warp_fwd_tc2|wgrad_tc|static_density_bwd|adam_flat|static_sigma_weight|composite|alpha_weight|time_embed|sample_points)'
timeout 900 ncu --set full --metrics lts__t_bytes.sum,l1tex__t_bytes.sum --clock-control none -k "$KR" --launch-skip 52 -c 26 -f -o $O/r02_step_full python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-cuda-eager > $O/r02_ncu_step.log 2>&1
python tools/kernel_table.py $O/r02_step_full.ncu-rep > $O/r02_kernel_table.json 2> $O/r02_kernel_table.err
python -c "
import json; d=json.load(open('$O/r02_kernel_table.json')); print(len(d['kernels']), d['total_ms_under_ncu'], list(d['kernels'])[:12])
" && cp $O/r02_kernel_table.json profiles/r02_kernel_table.json
rm -f $O/r02_step_full.ncu-rep
timeout 300 python -m pytest tests -m gpu -q --timeout 200 -k "saved_activations or two_tile or warp_forward_versions" > $O/r02_pytest_new.log 2>&1; echo "pytest rc $?" >> $O/r02_pytest_new.log; tail -3 $O/r02_pytest_new.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_train.json 2> $O/r02_bench_train.err
timeout 300 python bench.py --mode fwd --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_bench_fwd.json 2> $O/r02_bench_fwd.err
DYNRF_VARIANT=basis3 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02_bench_train_basis3.json 2> $O/r02_bench_train_basis3.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-graph --no-cpu-baseline --no-cuda-eager > $O/r02_bench_train_eager.json 2> $O/r02_bench_train_eager.err
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg3.json 2> $O/r02_bench_cfg3.err
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg4_pass1.json 2> $O/r02_bench_cfg4_pass1.err
timeout 900 python bench.py --config 2 --mode train9 --steps 5 --warmup 3 > $O/r02_bench_train9_300.json 2> $O/r02_bench_train9_300.err
timeout 900 python bench.py --config 4 --mode train9 --steps 5 --warmup 3 > $O/r02_bench_train9_640.json 2> $O/r02_bench_train9_640.err
timeout 900 python bench.py --config 5 --frames 1 > $O/r02_bench_render.json 2> $O/r02_bench_render.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*.json')):
    if 'gpu' in f.split('/')[-1].replace('gpurun','') : continue
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d.get('metric'), round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', (d.get('e2e') or {}).get('value'), 'eager_ref', (d.get('ref_cuda_eager') or {}).get('ms_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY
