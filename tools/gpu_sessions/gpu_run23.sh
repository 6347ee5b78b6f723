#!/bin/bash
# compute-before-wait in every operand-chunk producer (lookups / encodings first, operand-buffer wait after)
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > $O/r02m_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02m_pytest.log
tail -3 $O/r02m_pytest.log
timeout 200 python tools/exp_calls.py > $O/r02m_calls_new.txt 2> $O/r02m_calls_new.err
cut -c1-500 $O/r02m_calls_new.txt
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02m_train_new.json 2> $O/r02m_train_new.err
timeout 200 python bench.py --mode fwd --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02m_fwd_new.json 2> $O/r02m_fwd_new.err
python -c "
import json
for n in ('train','fwd'):
    d=json.loads(open('$O/r02m_%s_new.json' % n).read().strip().splitlines()[-1]); print(n, d['ms_per_step'], d['kernel_ms_per_step'])"
