#!/bin/bash
# timing experiments (wrong results on purpose): no weight TMA after the first chunks / no table gathers in the TC kernels
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
for V in "" base nowtma nogather; do
  DYNRF_VARIANT=$V timeout 200 python tools/exp_calls.py > $O/r02e_calls_${V:-new}.txt 2> $O/r02e_calls_${V:-new}.err
done
cat $O/r02e_calls_*.txt
