#!/bin/bash
# Round 2, GPU call U: Viterbi, one against two utterances per cluster (same kernel)
out=gpurun_out/r2u
mkdir -p $out
for pair in 0 1; do
PMN_VITERBI_PAIR=$pair timeout 600 python profiles/bench_preprocess.py --steps 3 --no-cpu > $out/pre_pair$pair.json 2> $out/pre_pair$pair.err; echo "rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/pre_pair$pair.json') if l.startswith('{')][-1])
print('pair $pair', d.get('ms_per_step'), {k: v for k, v in d.get('kernels', {}).items() if 'viterbi' in k})
PY
done
