#!/bin/bash
# A/B of the plane layout (z-fastest XZ / YZ planes vs round-1 x-fastest) on one box + the GPU suite on the new default
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -x > $O/r02b_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/r02b_pytest_gpu.log
for L in zfast xfast; do
  DYNRF_PLANE_LAYOUT=$L timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02b_train_$L.json 2> $O/r02b_train_$L.err
  DYNRF_PLANE_LAYOUT=$L timeout 300 python bench.py --mode fwd --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02b_fwd_$L.json 2> $O/r02b_fwd_$L.err
  DYNRF_PLANE_LAYOUT=$L timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --no-cuda-eager > $O/r02b_cfg3_$L.json 2> $O/r02b_cfg3_$L.err
  DYNRF_PLANE_LAYOUT=$L timeout 300 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline --no-cuda-eager > $O/r02b_cfg4_$L.json 2> $O/r02b_cfg4_$L.err
done
tail -3 $O/r02b_pytest_gpu.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b_*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), d.get('kernel_ms_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY
