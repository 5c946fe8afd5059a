#!/bin/bash
# Round 2, GPU call CC: train step time with the batched weight-norm backward
out=gpurun_out/r2cc
mkdir -p $out
for round in 1 2; do
timeout 600 python bench.py --only train --no-cpu-baseline > $out/bench_$round.json 2> $out/bench_$round.err; echo "rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/bench_$round.json') if l.startswith('{')][-1])
t = d['secondary']['train']
print(round(t['ms_per_step'], 3), t['value'], {k: v for k, v in sorted(t['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:8]})
PY
done
