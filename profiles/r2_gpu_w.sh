#!/bin/bash
# Round 2, GPU call W: cluster occupancy probe
out=gpurun_out/r2w
mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cluster_probe profiles/debug/cluster_probe.cu && /tmp/cluster_probe | tee $out/cluster_probe.txt
