#!/bin/bash
# Round 2, GPU call E: preprocess path after the Viterbi rewrite and the folded block 1
out=gpurun_out/r2e
mkdir -p $out
timeout 600 python -m pytest tests/test_preprocess_gpu.py tests/test_benchmark_shapes_gpu.py -q -x -k "not hifigan and not train" > $out/pytest.log 2>&1; echo "tests rc=$?"
tail -15 $out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --only preprocess > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2e/bench.json') if l.startswith('{')][-1])
p = d['secondary']['preprocess']
print(p['value'], p['ms_per_step'], p.get('parity'))
for k, v in sorted(p['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step']):
    print('  ', k, v)
PY
