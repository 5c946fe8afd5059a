#!/bin/bash
# Round 2, GPU call AE: DRAM traffic of every launch of one synthesis forward (new kernel mix),
# ncu --set full of the C = 32 and C = 64, k = 11 launches of conv1d_tcw_kernel
out=gpurun_out/r2ae
mkdir -p $out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -s 104 -c 110 --csv --log-file $out/forward_dram_traffic.csv python profiles/forward_once.py > $out/forward_once.log 2>&1; echo "traffic rc=$?"
tail -2 $out/forward_once.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tcw_kernel -c 1 \
    -o $out/tcw_c32_k11 -f python profiles/profile_tc_one.py 32 110080 11 c2 > $out/ncu_c32.log 2>&1; echo "ncu c32 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tcw_kernel -c 1 \
    -o $out/tcw_c64_k11 -f python profiles/profile_tc_one.py 64 55040 11 c2 > $out/ncu_c64.log 2>&1; echo "ncu c64 rc=$?"
