#!/bin/bash
# Round 2, GPU call AG: pitch block 1 with fp16 + 2 x fp8 operands
out=gpurun_out/r2ag
mkdir -p $out
timeout 900 python -m pytest tests/test_preprocess_gpu.py tests/test_benchmark_shapes_gpu.py tests/test_conv1d_tc_gpu.py -q --timeout 300 -k "not train" > $out/pytest.log 2>&1; echo "tests rc=$?"
tail -15 $out/pytest.log
for flag in 0 1; do
PMN_PITCH_F8=$flag timeout 600 python bench.py --steps 10 --warmup 3 --only preprocess > $out/bench_f8$flag.json 2> $out/bench_f8$flag.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/bench_f8$flag.json') if l.startswith('{')][-1])
p = d['secondary']['preprocess']
print('PMN_PITCH_F8=$flag', p['value'], p['ms_per_step'], p.get('parity'))
for k, v in sorted(p['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:4]:
    print('  ', k, v)
PY
done
