#!/bin/bash
# round-2 GPU session 1: tests, bench (both arms), K1 table, launch list, one full-step ncu capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_gpu.txt
nproc >> gpurun_out/r02_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_train_a.json 2> gpurun_out/r02_bench_train_a.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json ) 2> gpurun_out/r02_bench_reference.err
timeout 300 python tools/k1_table.py time > gpurun_out/r02_k1_time.json 2> gpurun_out/r02_k1_time.err
timeout 600 ncu --set full --clock-control none -k regex:k_static_sigma_weight -f -o gpurun_out/r02_k1 python tools/k1_table.py ncu > gpurun_out/r02_k1_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 700 --csv --log-file gpurun_out/r02_launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cuda-eager > gpurun_out/r02_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --launch-skip 900 -c 420 -f -o gpurun_out/r02_step_full python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-cuda-eager > gpurun_out/r02_ncu_step.log 2>&1
ls -la gpurun_out | tail -20
