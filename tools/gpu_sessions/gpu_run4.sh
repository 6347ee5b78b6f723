#!/bin/bash
# round-2 GPU session 4: full test suite, train bench (A/B of direct plane gradients), eager-API step, 9-pass step, render, fwd, launch list
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 --durations=8 > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/r02_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_train.json 2> $O/r02_bench_train.err
DYNRF_DIRECT_GRADS=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02_bench_train_nodirect.json 2> $O/r02_bench_train_nodirect.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-graph --no-cpu-baseline --no-cuda-eager > $O/r02_bench_train_eager.json 2> $O/r02_bench_train_eager.err
timeout 300 python bench.py --mode fwd --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02_bench_fwd.json 2> $O/r02_bench_fwd.err
timeout 900 python bench.py --config 2 --mode train9 --steps 5 --warmup 3 > $O/r02_bench_train9_300.json 2> $O/r02_bench_train9_300.err
timeout 900 python bench.py --config 4 --mode train9 --steps 5 --warmup 3 > $O/r02_bench_train9_640.json 2> $O/r02_bench_train9_640.err
timeout 900 python bench.py --config 5 --frames 1 > $O/r02_bench_render.json 2> $O/r02_bench_render.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 400 --csv --log-file $O/r02_launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cuda-eager > $O/r02_ncu_bench.log 2>&1
du -sh $O; ls -la $O | tail -12
