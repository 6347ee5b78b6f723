#!/bin/bash
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > $O/r02j_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02j_pytest.log
tail -6 $O/r02j_pytest.log
if ! grep -q "pytest rc 0" $O/r02j_pytest.log; then
  DYNRF_TC_WARP=1 timeout 600 python -m pytest tests -m gpu -q --timeout 300 > $O/r02j_pytest_warp1.log 2>&1; echo "pytest rc $?" >> $O/r02j_pytest_warp1.log
  tail -6 $O/r02j_pytest_warp1.log
fi
timeout 200 python tools/exp_calls.py > $O/r02j_calls_new.txt 2> $O/r02j_calls_new.err
cut -c1-600 $O/r02j_calls_new.txt
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02j_train_new.json 2> $O/r02j_train_new.err
timeout 200 python bench.py --mode fwd --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02j_fwd_new.json 2> $O/r02j_fwd_new.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02j_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), d.get('kernel_ms_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY
