#!/bin/bash
# round-2 GPU session 3 (after the alignment fix): tests, bench, launch list, per-kernel table, K1 table, cfg 4 / 5 lines.
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 --durations=10 > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/r02_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_train.json 2> $O/r02_bench_train.err
timeout 300 python tools/k1_table.py time > $O/r02_k1_time.json 2> $O/r02_k1_time.err
timeout 600 ncu --set full --metrics lts__t_bytes.sum,l1tex__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts.sum --clock-control none -k regex:k_static_sigma_weight -f -o $O/r02_k1 python tools/k1_table.py ncu > $O/r02_k1_ncu.log 2>&1
python tools/k1_table.py merge $O/r02_k1_time.json $O/r02_k1.ncu-rep > $O/r02_k1_table.json 2> $O/r02_k1_merge.err
rm -f $O/r02_k1.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 500 --csv --log-file $O/r02_launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cuda-eager > $O/r02_ncu_bench.log 2>&1
timeout 900 ncu --set full --metrics lts__t_bytes.sum,l1tex__t_bytes.sum --clock-control none -k 'regex:^k_' --launch-skip 120 -c 50 -f -o $O/r02_step_full python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-cuda-eager > $O/r02_ncu_step.log 2>&1
python tools/kernel_table.py $O/r02_step_full.ncu-rep > $O/r02_kernel_table.json 2> $O/r02_kernel_table.err
rm -f $O/r02_step_full.ncu-rep
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg4_pass1.json 2> $O/r02_bench_cfg4_pass1.err
timeout 900 python bench.py --config 2 --mode train9 --steps 5 --warmup 3 > $O/r02_bench_train9_300.json 2> $O/r02_bench_train9_300.err
timeout 900 python bench.py --config 4 --mode train9 --steps 5 --warmup 3 > $O/r02_bench_train9_640.json 2> $O/r02_bench_train9_640.err
timeout 900 python bench.py --config 5 --frames 1 > $O/r02_bench_render.json 2> $O/r02_bench_render.err
timeout 300 python bench.py --mode fwd --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > $O/r02_bench_fwd.json 2> $O/r02_bench_fwd.err
du -sh $O; ls -la $O | tail -30
