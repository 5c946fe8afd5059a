#!/bin/bash
# Round 2, GPU call BH: per-role cycle counters of the C = 256 / 128 launches, bf16 x 3 against the
# "fp16 + 2 x fp8" form
out=gpurun_out/r2bh
mkdir -p $out
for c in 256 128; do
PMN_TCW=0 timeout 300 python profiles/tc_breakdown.py $c > $out/breakdown_${c}_bf16.txt 2>&1; echo "rc=$?"
PMN_TCW=0 timeout 300 python profiles/tc_breakdown.py $c f8 > $out/breakdown_${c}_f8.txt 2>&1; echo "rc=$?"
cat $out/breakdown_${c}_bf16.txt $out/breakdown_${c}_f8.txt
done
