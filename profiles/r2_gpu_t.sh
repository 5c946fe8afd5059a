#!/bin/bash
# Round 2, GPU call T: Viterbi with two utterances per cluster frame-major blocks 3-5 + head of the pitch network, Viterbi score loads in groups of 8
out=gpurun_out/r2t
mkdir -p $out
timeout 900 python -m pytest tests/test_preprocess_gpu.py tests/test_benchmark_shapes_gpu.py -q -x --timeout 300 -k "not train" > $out/pytest.log 2>&1; echo "tests rc=$?"
tail -4 $out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --only preprocess > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2t/bench.json') if l.startswith('{')][-1])
p = d['secondary']['preprocess']
print(p['value'], p['ms_per_step'], p.get('parity'))
for k, v in sorted(p['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:12]:
    print('  ', k, v)
PY
