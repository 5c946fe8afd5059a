#!/bin/bash
# Round 2, GPU call AR: where a FARGAN subframe's cycles go (waiting for operands against local work)
out=gpurun_out/r2ar
mkdir -p $out
PMN_FARGAN_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 3 --only fargan > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
grep "fargan CTA 0" $out/bench.err | tail -2
