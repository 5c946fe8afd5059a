#!/bin/bash
# Round 2, GPU call AK: Viterbi scan as two independent compare chains
out=gpurun_out/r2ak
mkdir -p $out
timeout 900 python -m pytest tests/test_preprocess_gpu.py tests/test_benchmark_shapes_gpu.py -q --timeout 300 -k "not train" > $out/pytest.log 2>&1; echo "tests rc=$?"
tail -2 $out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --only preprocess > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/bench.json') if l.startswith('{')][-1])
p = d['secondary']['preprocess']
print(p['value'], p['ms_per_step'], p.get('parity'))
for k, v in sorted(p['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:4]:
    print('  ', k, v)
PY
