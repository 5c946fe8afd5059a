#!/bin/bash
# Round 2, GPU call BR: pair masks compared by the bench itself (two alternating runs each)
out=gpurun_out/r2br
mkdir -p $out
for round in 1 2 3; do
for mask in 0x008 0x208 0x608 0x200; do
  PMN_PAIR_MASK=$mask timeout 600 python bench.py --no-secondary --no-cpu-baseline > $out/bench_${mask}_$round.json 2> $out/bench_${mask}_$round.err
  python - <<PY
import json
d = json.loads([l for l in open('$out/bench_${mask}_$round.json') if l.startswith('{')][-1])
k = d['roofline']['kernels']
print('$mask run $round', round(d['ms_per_step'], 3), {n: round(k[n]['ms_per_step'], 3) for n in ('conv1d_tc_kernel', 'conv1d_tcw_kernel', 'conv_pair_tc_kernel')})
PY
done
done
