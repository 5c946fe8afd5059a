#!/bin/bash
# Round 2, GPU call BJ: conv1d_tc_kernel with three warpgroups + setmaxnreg and the rare epilogue options
# compiled per variant: tests, per-role breakdown, step time with and without the fp8 form
out=gpurun_out/r2bj
mkdir -p $out
timeout 900 python -m pytest tests/test_conv1d_tc_gpu.py tests/test_generator_gpu.py tests/test_conv_pair_tc_gpu.py -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
timeout 600 python -m pytest tests/test_preprocess_gpu.py tests/test_benchmark_shapes_gpu.py -q -x > $out/pytest2.log 2>&1; echo "pytest2 rc=$?"; tail -3 $out/pytest2.log
PMN_TCW=0 timeout 300 python profiles/tc_breakdown.py > $out/breakdown_bf16.txt 2>&1; echo "rc=$?"
grep -E "k= 3|k=11 c1|k=11 c2 " $out/breakdown_bf16.txt | cut -c1-250
for c in 256 128; do PMN_TCW=0 timeout 300 python profiles/tc_breakdown.py $c f8 > $out/breakdown_${c}_f8.txt 2>&1; cat $out/breakdown_${c}_f8.txt | cut -c1-250; done
for f8 in 0 1; do
PMN_GENERATOR_F8=$f8 timeout 600 python bench.py --no-secondary --no-cpu-baseline > $out/bench_f8_$f8.json 2> $out/bench_f8_$f8.err; echo "bench f8=$f8 rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/bench_f8_$f8.json') if l.startswith('{')][-1])
print('f8=$f8', d['ms_per_step'], d['value'], 'frac', d['roofline']['frac'])
for k, v in sorted(d['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:4]: print('  ', k, v)
PY
done
for f8 in 0 1; do
PMN_PITCH_F8=$f8 timeout 600 python profiles/bench_preprocess.py --steps 5 --no-cpu > $out/preprocess_f8_$f8.json 2> $out/preprocess_f8_$f8.err; echo "preprocess f8=$f8 rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/preprocess_f8_$f8.json') if l.startswith('{')][-1])
print('pitch f8=$f8', d['ms_per_step'], {k: v['ms'] for k, v in d['kernels'].items() if k in ('conv1d_tc_kernel', 'shared_norm_planes_kernel')})
PY
done
