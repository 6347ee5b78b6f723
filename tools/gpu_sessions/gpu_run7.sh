#!/bin/bash
# round-2 final 1-GPU session: full suite, smoke, every bench line, launch list, per-kernel table
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 --durations=5 > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/r02_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_train.json 2> $O/r02_bench_train.err
timeout 300 python bench.py --mode fwd --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_bench_fwd.json 2> $O/r02_bench_fwd.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-graph --no-cpu-baseline --no-cuda-eager > $O/r02_bench_train_eager.json 2> $O/r02_bench_train_eager.err
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg3.json 2> $O/r02_bench_cfg3.err
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg4_pass1.json 2> $O/r02_bench_cfg4_pass1.err
timeout 900 python bench.py --config 2 --mode train9 --steps 5 --warmup 3 > $O/r02_bench_train9_300.json 2> $O/r02_bench_train9_300.err
timeout 900 python bench.py --config 4 --mode train9 --steps 5 --warmup 3 > $O/r02_bench_train9_640.json 2> $O/r02_bench_train9_640.err
timeout 900 python bench.py --config 5 --frames 1 > $O/r02_bench_render.json 2> $O/r02_bench_render.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 400 --csv --log-file $O/r02_launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cuda-eager > $O/r02_ncu_bench.log 2>&1
timeout 900 ncu --set full --metrics lts__t_bytes.sum,l1tex__t_bytes.sum --clock-control none -k 'regex:^k_' --launch-skip 120 -c 50 -f -o $O/r02_step_full python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-cuda-eager > $O/r02_ncu_step.log 2>&1
python tools/kernel_table.py $O/r02_step_full.ncu-rep > $O/r02_kernel_table.json 2> $O/r02_kernel_table.err
rm -f $O/r02_step_full.ncu-rep
du -sh $O
