#!/bin/bash
# final commit: the two headline bench lines + smoke
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/r02_smoke.log 2>&1; tail -1 $O/r02_smoke.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_train.json 2> $O/r02_bench_train.err
timeout 300 python bench.py --mode fwd --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_bench_fwd.json 2> $O/r02_bench_fwd.err
python -c "
import json
for n in ('train','fwd'):
    d=json.loads(open('$O/r02_bench_%s.json' % n).read().strip().splitlines()[-1]); print(n, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('ref_cuda_eager',{}).get('ms_per_step'))"
