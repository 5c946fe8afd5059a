#!/bin/bash
# Round 2, GPU call BD: per-role cycle breakdown of conv1d_tc_kernel at C = 32 (time on the M side)
out=gpurun_out/r2bd
mkdir -p $out
PMN_TCW=0 timeout 300 python profiles/tc_breakdown.py 32 | tee $out/tc_breakdown_c32.txt
