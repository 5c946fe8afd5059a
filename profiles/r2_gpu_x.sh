#!/bin/bash
# Round 2, GPU call X: Viterbi with 1 .. 4 utterances per cluster, then tests + bench with the automatic choice
out=gpurun_out/r2x
mkdir -p $out
for pair in 1 2 3 4; do
PMN_VITERBI_DEBUG=1 PMN_VITERBI_PAIR=$pair timeout 600 python profiles/bench_preprocess.py --steps 3 --no-cpu > $out/pre_pair$pair.json 2> $out/pre_pair$pair.err; echo "rc=$?"
grep "viterbi CTA" $out/pre_pair$pair.err | tail -1
python - <<PY
import json
d = json.loads([l for l in open('$out/pre_pair$pair.json') if l.startswith('{')][-1])
print('per cluster $pair', d.get('ms_per_step'), {k: v for k, v in d.get('kernels', {}).items() if 'viterbi_cluster' in k})
PY
done
timeout 900 python -m pytest tests/test_preprocess_gpu.py tests/test_benchmark_shapes_gpu.py -q -x --timeout 300 -k "not train" > $out/pytest.log 2>&1; echo "tests rc=$?"
tail -3 $out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --only preprocess > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2x/bench.json') if l.startswith('{')][-1])
p = d['secondary']['preprocess']
print(p['value'], p['ms_per_step'], p.get('parity'))
for k, v in sorted(p['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:4]:
    print('  ', k, v)
PY
