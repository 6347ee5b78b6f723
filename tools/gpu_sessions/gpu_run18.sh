#!/bin/bash
# final build, pass 1: suite, smoke, launch list + per-kernel ncu table of one train step (bench.py reads the table from profiles/)
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --durations=5 > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/r02_pytest_gpu.log
tail -3 $O/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/r02_smoke.log 2>&1; tail -2 $O/r02_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 400 --csv --log-file $O/r02_launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cuda-eager > $O/r02_ncu_bench.log 2>&1
timeout 900 ncu --set full --metrics lts__t_bytes.sum,l1tex__t_bytes.sum --clock-control none -k 'regex:^k_' --launch-skip 120 -c 50 -f -o $O/r02_step_full python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-cuda-eager > $O/r02_ncu_step.log 2>&1
python tools/kernel_table.py $O/r02_step_full.ncu-rep > $O/r02_kernel_table.json 2> $O/r02_kernel_table.err
rm -f $O/r02_step_full.ncu-rep
head -c 600 $O/r02_kernel_table.json
