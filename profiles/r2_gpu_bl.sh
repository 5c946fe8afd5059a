#!/bin/bash
# Round 2, GPU call BL: residual pipeline across tiles + chunk loop unrolled in groups: tests, breakdown,
# step time
out=gpurun_out/r2bl
mkdir -p $out
timeout 900 python -m pytest tests/test_conv1d_tc_gpu.py tests/test_generator_gpu.py tests/test_conv_pair_tc_gpu.py tests/test_benchmark_shapes_gpu.py tests/test_preprocess_gpu.py -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
for f8 in 1 0; do
PMN_GENERATOR_F8=$f8 timeout 600 python bench.py --no-secondary --no-cpu-baseline > $out/bench_f8_$f8.json 2> $out/bench_f8_$f8.err; echo "bench f8=$f8 rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/bench_f8_$f8.json') if l.startswith('{')][-1])
print('f8=$f8', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
for k, v in sorted(d['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:4]: print('  ', k, v)
PY
done
PMN_TCW=0 timeout 300 python profiles/tc_breakdown.py 128 f8 > $out/breakdown_128_f8.txt 2>&1; cut -c1-250 $out/breakdown_128_f8.txt
PMN_TCW=0 timeout 300 python profiles/tc_breakdown.py > $out/breakdown_bf16.txt 2>&1; grep -E "C=256|C= 32|C=128 k= 3" $out/breakdown_bf16.txt | cut -c1-250
timeout 600 python profiles/bench_preprocess.py --steps 5 --no-cpu > $out/preprocess.json 2> $out/preprocess.err; echo "preprocess rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/preprocess.json') if l.startswith('{')][-1])
print('preprocess', d['ms_per_step'], {k: v['ms'] for k, v in d['kernels'].items() if k in ('conv1d_tc_kernel', 'shared_norm_planes_kernel')})
PY
