#!/bin/bash
# Round 2, GPU call BW: why is the fp8-form operand writer of pitch block 1 twice as slow as the bf16 one?
out=gpurun_out/r2bw
mkdir -p $out
PMN_PITCH_F8=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:shared_norm_planes_kernel -c 1 \
    -o $out/shared_norm_f8 -f python profiles/bench_preprocess.py --steps 1 --no-cpu > $out/ncu_a.log 2>&1; echo "rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:shared_norm_planes_kernel -c 1 \
    -o $out/shared_norm_bf16 -f python profiles/bench_preprocess.py --steps 1 --no-cpu > $out/ncu_b.log 2>&1; echo "rc=$?"
