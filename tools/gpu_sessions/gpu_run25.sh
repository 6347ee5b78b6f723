#!/bin/bash
# K1 table (north-star kernel) with the shipped plane layout (z-fastest XZ / YZ planes)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/k1_table.py time > $O/r02_k1_time.json 2> $O/r02_k1_time.err
timeout 600 ncu --set full --metrics lts__t_bytes.sum,l1tex__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts.sum --clock-control none -k regex:k_static_sigma_weight -f -o $O/r02_k1 python tools/k1_table.py ncu > $O/r02_k1_ncu.log 2>&1
python tools/k1_table.py merge $O/r02_k1_time.json $O/r02_k1.ncu-rep > $O/r02_k1_table.json 2> $O/r02_k1_merge.err
rm -f $O/r02_k1.ncu-rep
head -c 1500 $O/r02_k1_table.json
