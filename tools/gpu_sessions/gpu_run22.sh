#!/bin/bash
# head backward: E4 merged with the next tile's S (4-stage alternation)
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > $O/r02l_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02l_pytest.log
tail -3 $O/r02l_pytest.log
timeout 200 python tools/exp_calls.py > $O/r02l_calls_new.txt 2> $O/r02l_calls_new.err
cut -c1-500 $O/r02l_calls_new.txt
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02l_train_new.json 2> $O/r02l_train_new.err
python -c "
import json; d=json.loads(open('$O/r02l_train_new.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernel_ms_per_step'])"
