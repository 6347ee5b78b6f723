#!/bin/bash
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x > $O/r02i_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02i_pytest.log
tail -4 $O/r02i_pytest.log
timeout 200 python tools/exp_calls.py > $O/r02i_calls_new.txt 2> $O/r02i_calls_new.err
cut -c1-600 $O/r02i_calls_new.txt
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02i_train_new.json 2> $O/r02i_train_new.err
timeout 200 python bench.py --mode fwd --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02i_fwd_new.json 2> $O/r02i_fwd_new.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02i_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), d.get('kernel_ms_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY
