#!/bin/bash
# Round 2, GPU call BB: ConvTranspose epilogue writes the planes of its output
out=gpurun_out/r2bb
mkdir -p $out
timeout 900 python -m pytest tests/test_generator_gpu.py tests/test_benchmark_shapes_gpu.py tests/test_conv1d_tc_gpu.py tests/test_conv_pair_tc_gpu.py -q --timeout 300 -k "not train and not preprocess" > $out/pytest.log 2>&1; echo "tests rc=$?"
tail -3 $out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --only synthesis > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2bb/bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['parity']['max_rel_err'], d['gpu_launches'])
for k, v in sorted(d['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:8]:
    print('  ', k, v)
PY
