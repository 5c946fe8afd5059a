#!/bin/bash
# Round 2, GPU call V: where the Viterbi kernel's time goes (prologue / frame loop / backtrace)
out=gpurun_out/r2v
mkdir -p $out
for pair in 0 1; do
PMN_VITERBI_DEBUG=1 PMN_VITERBI_PAIR=$pair timeout 600 python profiles/bench_preprocess.py --steps 1 --no-cpu > $out/pre_pair$pair.json 2> $out/pre_pair$pair.err; echo "rc=$?"
grep "viterbi CTA" $out/pre_pair$pair.err | tail -2
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm --format=csv
