#!/bin/bash
# Round 2, GPU call L: packed bf16 converts, flat LayerNorm statistics, shared-norm rewrite,
# Viterbi publishing 4 states per 16-byte st.async
out=gpurun_out/r2l
mkdir -p $out
timeout 900 python -m pytest tests/test_preprocess_gpu.py tests/test_benchmark_shapes_gpu.py tests/test_conv1d_tc_gpu.py tests/test_conv_pair_tc_gpu.py tests/test_generator_gpu.py -q -x --timeout 300 -k "not train" > $out/pytest.log 2>&1; echo "tests rc=$?"
tail -8 $out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --only preprocess > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
tail -c 300 $out/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2l/bench.json') if l.startswith('{')][-1])
print('synthesis', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['parity'])
p = d['secondary']['preprocess']
print(p['value'], p['ms_per_step'], p.get('parity'))
for k, v in sorted(p['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:10]:
    print('  ', k, v)
for k, v in sorted(d['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:10]:
    print('  ', k, v)
PY
