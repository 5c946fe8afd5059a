#!/bin/bash
# Round 2, GPU call CA: ncu --set full of one launch of each kind on the shipped synthesis path
out=gpurun_out/r2ca
mkdir -p $out
cap() {  # name kernel-regex args...
  name=$1; regex=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex -c 1 -o $out/$name -f "$@" > $out/$name.log 2>&1; echo "$name rc=$?"
}
cap tc_c256_k11_c1 conv1d_tc_kernel python profiles/profile_tc_one.py 256 3440 11 c1
cap tc_c128_k11_c1_f8 conv1d_tc_kernel python profiles/profile_tc_one.py 128 27520 11 c1 f8
cap tc_c128_k7_c2_f8 conv1d_tc_kernel python profiles/profile_tc_one.py 128 27520 7 c2 f8
cap tcw_c64_k11_c1 conv1d_tcw_kernel python profiles/profile_tc_one.py 64 55040 11 c1
cap tcw_c32_k11_c2 conv1d_tcw_kernel python profiles/profile_tc_one.py 32 110080 11 c2
cap tc_c64_k7_c2 conv1d_tc_kernel python profiles/profile_tc_one.py 64 55040 7 c2
cap tc_c32_k3_c1 conv1d_tc_kernel python profiles/profile_tc_one.py 32 110080 3 c1
