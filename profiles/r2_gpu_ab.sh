#!/bin/bash
# Round 2, GPU call AB: per-layer comparison of the two narrow-layer kernels
out=gpurun_out/r2ab
mkdir -p $out
timeout 900 python profiles/narrow_layers.py | tee $out/narrow_layers.txt
