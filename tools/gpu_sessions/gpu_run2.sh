#!/bin/bash
# round-2 GPU session: tests, bench (both arms + other configs), K1 table, launch list, per-kernel table of one train step.
# gpurun copies back at most 64 MiB of gpurun_out/: every .ncu-rep is reduced to CSV / JSON on the box and deleted.
set -x
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r02_gpu.txt
nproc >> $O/r02_gpu.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 --durations=15 > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/r02_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_train.json 2> $O/r02_bench_train.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference.json ) 2> $O/r02_bench_reference.err
timeout 300 python tools/k1_table.py time > $O/r02_k1_time.json 2> $O/r02_k1_time.err
timeout 600 ncu --set full --clock-control none -k regex:k_static_sigma_weight -f -o $O/r02_k1 python tools/k1_table.py ncu > $O/r02_k1_ncu.log 2>&1
python tools/k1_table.py merge $O/r02_k1_time.json $O/r02_k1.ncu-rep > $O/r02_k1_table.json 2> $O/r02_k1_merge.err
rm -f $O/r02_k1.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 500 -c 600 --csv --log-file $O/r02_launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cuda-eager > $O/r02_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none -k 'regex:^k_' --launch-skip 150 -c 60 -f -o $O/r02_step_full python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-cuda-eager > $O/r02_ncu_step.log 2>&1
python tools/kernel_table.py $O/r02_step_full.ncu-rep > $O/r02_kernel_table.json 2> $O/r02_kernel_table.err
rm -f $O/r02_step_full.ncu-rep
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg3.json 2> $O/r02_bench_cfg3.err
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline --no-cuda-eager > $O/r02_bench_cfg4_pass1.json 2> $O/r02_bench_cfg4_pass1.err
timeout 900 python bench.py --config 2 --mode train9 --steps 5 --warmup 3 > $O/r02_bench_train9_300.json 2> $O/r02_bench_train9_300.err
timeout 900 python bench.py --config 4 --mode train9 --steps 5 --warmup 3 > $O/r02_bench_train9_640.json 2> $O/r02_bench_train9_640.err
timeout 900 python bench.py --config 5 --frames 1 > $O/r02_bench_render.json 2> $O/r02_bench_render.err
du -sh $O; ls -la $O | tail -30
