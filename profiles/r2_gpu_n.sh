#!/bin/bash
# Round 2, GPU call O: Viterbi band in registers
out=gpurun_out/r2o
mkdir -p $out
PMN_VITERBI_DEBUG=1 timeout 600 python profiles/bench_preprocess.py --steps 1 --no-cpu > $out/viterbi_debug.json 2> $out/viterbi_debug.err; echo "rc=$?"
grep "viterbi rank" $out/viterbi_debug.err | tail -16
