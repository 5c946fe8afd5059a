#!/bin/bash
# Round 2, GPU call K: launch list of one preprocess step, ncu --set full of the new block-0
# LayerNorm kernel and of the narrow residual-block launches (C = 32 / 64, k = 11)
out=gpurun_out/r2k
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file $out/launches_pre.csv python profiles/bench_preprocess.py --steps 1 --no-cpu > $out/launches_pre.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'shared_norm_planes_kernel|column_sums_kernel' --launch-skip 0 -c 3 -o $out/shared_norm -f \
    python profiles/bench_preprocess.py --steps 1 --no-cpu > $out/ncu_norm.log 2>&1; echo "ncu norm rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_kernel -c 1 \
    -o $out/tc_c32_k11 -f python profiles/profile_tc_one.py 32 110080 11 c2 > $out/ncu_c32.log 2>&1; echo "ncu c32 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_kernel -c 1 \
    -o $out/tc_c64_k11 -f python profiles/profile_tc_one.py 64 55040 11 c2 > $out/ncu_c64.log 2>&1; echo "ncu c64 rc=$?"
ls -la $out
