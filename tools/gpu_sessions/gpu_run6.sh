#!/bin/bash
# A/B on ONE box, identical model: round-1 tree (_r1/) vs HEAD (stream hints on) ; A/B of direct plane gradients
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
for i in 1 2; do
  (cd _r1 && timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_ab2_r1_$i.json 2> $O/r02_ab2_r1_$i.err)
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02_ab2_head_$i.json 2> $O/r02_ab2_head_$i.err
  DYNRF_DIRECT_GRADS=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02_ab2_nodirect_$i.json 2> $O/r02_ab2_nodirect_$i.err
done
(cd _r1 && timeout 300 python bench.py --mode fwd --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_ab2_r1_fwd.json 2> $O/r02_ab2_r1_fwd.err)
timeout 300 python bench.py --mode fwd --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02_ab2_head_fwd.json 2> $O/r02_ab2_head_fwd.err
ls -la $O | tail -5
