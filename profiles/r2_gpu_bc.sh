#!/bin/bash
# Round 2, GPU call BC: ncu --set full of the C = 32 launches that sit far from their HBM floor
out=gpurun_out/r2bc
mkdir -p $out
PMN_TCW=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_kernel -c 1 \
    -o $out/tc_c32_k3_c1 -f python profiles/profile_tc_one.py 32 110080 3 c1 > $out/ncu_a.log 2>&1; echo "rc=$?"
PMN_TCW=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_kernel -c 1 \
    -o $out/tc_c32_k3_c2 -f python profiles/profile_tc_one.py 32 110080 3 c2 > $out/ncu_b.log 2>&1; echo "rc=$?"
