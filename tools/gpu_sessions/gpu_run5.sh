#!/bin/bash
# A/B on ONE box: round-1 tree (_r1/) vs HEAD vs HEAD + streaming cache hints; plus the gradient tests
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
timeout 600 python -m pytest tests/test_gpu_grad.py tests/test_gpu_step.py -m gpu -q --timeout 600 > $O/r02_pytest_grad.log 2>&1; echo "pytest rc $?" >> $O/r02_pytest_grad.log
for i in 1 2; do
  (cd _r1 && timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_ab_r1_$i.json 2> $O/r02_ab_r1_$i.err)
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02_ab_head_$i.json 2> $O/r02_ab_head_$i.err
  DYNRF_VARIANT=stream timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02_ab_stream_$i.json 2> $O/r02_ab_stream_$i.err
done
DYNRF_VARIANT=stream timeout 300 python bench.py --mode fwd --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02_ab_stream_fwd.json 2> $O/r02_ab_stream_fwd.err
DYNRF_VARIANT=stream timeout 600 python -m pytest tests/test_gpu_grad.py tests/test_gpu_parity.py -m gpu -q --timeout 600 > $O/r02_pytest_stream.log 2>&1; echo "pytest rc $?" >> $O/r02_pytest_stream.log
ls -la $O | tail -12
