#!/bin/bash
# Round 2, GPU call H: st.async Viterbi + staged backtrace, pipelined host streaming
out=gpurun_out/r2h
mkdir -p $out
timeout 600 python -m pytest tests/test_preprocess_gpu.py tests/test_benchmark_shapes_gpu.py tests/test_generator_gpu.py -q -x --timeout 120 -k "not train" > $out/pytest.log 2>&1; echo "tests rc=$?"
tail -15 $out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --only preprocess > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
tail -c 400 $out/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2h/bench.json') if l.startswith('{')][-1])
print('synthesis', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'sync', d['e2e']['synchronous_value'])
p = d['secondary']['preprocess']
print(p['value'], p['ms_per_step'], p.get('parity'))
for k, v in sorted(p['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:5]:
    print('  ', k, v)
PY
