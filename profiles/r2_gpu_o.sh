#!/bin/bash
# Round 2, GPU call O: Viterbi band in registers
out=gpurun_out/r2o
mkdir -p $out
PMN_VITERBI_DEBUG=1 timeout 600 python profiles/bench_preprocess.py --steps 1 --no-cpu > $out/viterbi_debug.json 2> $out/viterbi_debug.err; echo "rc=$?"
grep "viterbi rank" $out/viterbi_debug.err | tail -16
timeout 900 python -m pytest tests/test_preprocess_gpu.py tests/test_benchmark_shapes_gpu.py -q -x --timeout 300 -k "not train" > $out/pytest.log 2>&1; echo "tests rc=$?"
tail -3 $out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --only preprocess > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2o/bench.json') if l.startswith('{')][-1])
p = d['secondary']['preprocess']
print(p['value'], p['ms_per_step'], p.get('parity'))
for k, v in sorted(p['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:6]:
    print('  ', k, v)
PY
