#!/bin/bash
# multi-GPU session (N = number of visible GPUs): NCCL sharded == unsharded test, weak + strong scaling of the train step,
# the 9-pass step and the render path
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 900 -s > $O/r02_pytest_multi_${N}gpu.log 2>&1; echo "pytest rc $?" >> $O/r02_pytest_multi_${N}gpu.log
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_bench_train_${N}gpu.json 2> $O/r02_bench_train_${N}gpu.err
DYNRF_STRONG=1 timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_bench_train_${N}gpu_strong.json 2> $O/r02_bench_train_${N}gpu_strong.err
timeout 900 $TR bench.py --gpus $N --config 2 --mode train9 --steps 5 --warmup 3 > $O/r02_bench_train9_300_${N}gpu.json 2> $O/r02_bench_train9_300_${N}gpu.err
DYNRF_STRONG=1 timeout 900 $TR bench.py --gpus $N --config 2 --mode train9 --steps 5 --warmup 3 > $O/r02_bench_train9_300_${N}gpu_strong.json 2> $O/r02_bench_train9_300_${N}gpu_strong.err
timeout 900 $TR bench.py --gpus $N --config 5 --frames 1 > $O/r02_bench_render_${N}gpu.json 2> $O/r02_bench_render_${N}gpu.err
ls -la $O | tail -14
