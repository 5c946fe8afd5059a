#!/bin/bash
# Round 2, GPU call AA: narrow layers with the weights on the M side (conv1d_tcw_kernel)
out=gpurun_out/r2aa
mkdir -p $out
timeout 900 python -m pytest tests/test_conv1d_tc_gpu.py tests/test_generator_gpu.py tests/test_benchmark_shapes_gpu.py -q -x --timeout 300 -k "not train and not preprocess" > $out/pytest.log 2>&1; echo "tests rc=$?"
tail -12 $out/pytest.log
for flag in 0 1; do
PMN_TCW=$flag timeout 600 python bench.py --steps 10 --warmup 3 --only synthesis > $out/bench_tcw$flag.json 2> $out/bench_tcw$flag.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/bench_tcw$flag.json') if l.startswith('{')][-1])
print('PMN_TCW=$flag', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['parity']['max_rel_err'])
for k, v in sorted(d['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:5]:
    print('  ', k, v)
PY
done
