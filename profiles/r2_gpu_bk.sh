#!/bin/bash
# Round 2, GPU call BK: fp8 form at C = 128 only (default on), accumulate input prefetched: tests, step
# time, ncu --set full of an epilogue-bound launch of each kind
out=gpurun_out/r2bk
mkdir -p $out
timeout 900 python -m pytest tests/test_conv1d_tc_gpu.py tests/test_generator_gpu.py tests/test_conv_pair_tc_gpu.py tests/test_benchmark_shapes_gpu.py tests/test_synthesize_gpu.py -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
for f8 in 1 0; do
PMN_GENERATOR_F8=$f8 timeout 600 python bench.py --no-secondary --no-cpu-baseline > $out/bench_f8_$f8.json 2> $out/bench_f8_$f8.err; echo "bench f8=$f8 rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/bench_f8_$f8.json') if l.startswith('{')][-1])
print('f8=$f8', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'parity', d['parity'])
for k, v in sorted(d['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:4]: print('  ', k, v)
PY
done
PMN_TCW=0 timeout 300 python profiles/tc_breakdown.py 128 f8 > $out/breakdown_128_f8.txt 2>&1; cut -c1-250 $out/breakdown_128_f8.txt
PMN_TCW=0 timeout 300 python profiles/tc_breakdown.py 32 > $out/breakdown_32.txt 2>&1; cut -c1-250 $out/breakdown_32.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_kernel -c 1 \
    -o $out/tc_c128_k7_c2_f8 -f python profiles/profile_tc_one.py 128 27520 7 c2 f8 > $out/ncu_a.log 2>&1; echo "rc=$?"
PMN_TCW=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_kernel -c 1 \
    -o $out/tc_c32_k3_c2 -f python profiles/profile_tc_one.py 32 110080 3 c2 > $out/ncu_b.log 2>&1; echo "rc=$?"
