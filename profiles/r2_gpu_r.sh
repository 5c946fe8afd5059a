#!/bin/bash
# Round 2, GPU call R: launch list of one preprocess step after the frame-major layers
out=gpurun_out/r2r
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file $out/launches_pre.csv python profiles/bench_preprocess.py --steps 1 --no-cpu > $out/launches_pre.log 2>&1; echo "launch list rc=$?"
