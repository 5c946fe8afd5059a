#!/bin/bash
# Round 2, GPU call AF: DRAM traffic of every launch of one synthesis forward (new kernel mix),
# ncu --set full of the C = 32 and C = 64, k = 11 launches of conv1d_tcw_kernel
out=gpurun_out/r2af
mkdir -p $out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -c 400 --csv --log-file $out/forward_dram_traffic.csv python profiles/forward_once.py > $out/forward_once.log 2>&1; echo "traffic rc=$?"
tail -2 $out/forward_once.log
