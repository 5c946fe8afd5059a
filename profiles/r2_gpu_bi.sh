#!/bin/bash
# Round 2, GPU call BI: which commit slowed the conv1d_tc epilogue (C = 128 / 32, k = 3)? cycle breakdown
# of four earlier trees against HEAD
out=gpurun_out/r2bi
mkdir -p $out
root=$PWD
for sha in 85ce8d7 c5cf8e1 9a88f8c 5fa091d; do
  cd $root/profiles/debug/bisect/$sha
  PMN_TCW=0 timeout 300 python profiles/tc_breakdown.py > $root/$out/$sha.txt 2>&1; echo "$sha rc=$?"
  grep -E "C=128 k= 3|C= 32 k= 3|C=128 k=11 c1" $root/$out/$sha.txt | cut -c1-250
done
cd $root
PMN_TCW=0 timeout 300 python profiles/tc_breakdown.py > $out/head.txt 2>&1; echo "head rc=$?"
grep -E "C=128 k= 3|C= 32 k= 3|C=128 k=11 c1" $out/head.txt | cut -c1-250
