#!/bin/bash
# Round 2, GPU call J: shared block 0 with its own LayerNorm kernels (column sums, frame statistics, 8 frames x 8 channels per CTA)
out=gpurun_out/r2j
mkdir -p $out
timeout 900 python -m pytest tests/test_preprocess_gpu.py tests/test_benchmark_shapes_gpu.py -q -x --timeout 300 -k "not train" > $out/pytest.log 2>&1; echo "tests rc=$?"
tail -15 $out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --only preprocess > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
tail -c 400 $out/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2j/bench.json') if l.startswith('{')][-1])
print('synthesis', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
p = d['secondary']['preprocess']
print(p['value'], p['ms_per_step'], p.get('parity'))
for k, v in sorted(p['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:8]:
    print('  ', k, v)
PY
