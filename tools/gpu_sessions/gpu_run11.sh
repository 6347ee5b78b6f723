#!/bin/bash
# dual-tile k_app_fwd_tc: parity suite + forward / train bench against the previous build (variant "base")
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_paths.py -m gpu -q --timeout 120 -x > $O/r02d_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02d_pytest.log
tail -5 $O/r02d_pytest.log
for V in "" base; do
  DYNRF_VARIANT=$V timeout 200 python bench.py --mode fwd --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02d_fwd_${V:-new}.json 2> $O/r02d_fwd_${V:-new}.err
  DYNRF_VARIANT=$V timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02d_train_${V:-new}.json 2> $O/r02d_train_${V:-new}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02d_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), d.get('kernel_ms_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY
