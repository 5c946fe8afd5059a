#!/bin/bash
# Round 2, GPU call BO: epilogue address arithmetic by pointer increments (half of the epilogue's
# instructions were 64-bit address computations): tests, same-box A/B against the commit before
out=gpurun_out/r2bo
mkdir -p $out
root=$PWD
timeout 900 python -m pytest tests/test_conv1d_tc_gpu.py tests/test_generator_gpu.py tests/test_conv_pair_tc_gpu.py tests/test_benchmark_shapes_gpu.py tests/test_preprocess_gpu.py -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
run() {  # tree label
  cd $1
  timeout 600 python bench.py --no-secondary --no-cpu-baseline > $root/$out/bench_$2.json 2> $root/$out/bench_$2.err
  cd $root
  python - <<PY
import json
d = json.loads([l for l in open('$out/bench_$2.json') if l.startswith('{')][-1])
k = d['roofline']['kernels']
print('$2', round(d['ms_per_step'], 3), {n: round(k[n]['ms_per_step'], 3) for n in ('conv1d_tc_kernel', 'conv1d_tcw_kernel', 'conv_pair_tc_kernel', 'conv_transpose1d_tc_kernel')})
PY
}
for round in 1 2; do
  run $root/profiles/debug/ab/before before$round
  run $root head$round
done
PMN_TCW=0 timeout 300 python profiles/tc_breakdown.py > $out/breakdown_bf16.txt 2>&1; grep -E "k= 3|k=11 c2 " $out/breakdown_bf16.txt | cut -c1-250
PMN_TCW=0 timeout 300 python profiles/tc_breakdown.py 128 f8 > $out/breakdown_128_f8.txt 2>&1; cut -c1-250 $out/breakdown_128_f8.txt
timeout 600 python profiles/bench_preprocess.py --steps 5 --no-cpu > $out/preprocess.json 2> $out/preprocess.err; echo "preprocess rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/preprocess.json') if l.startswith('{')][-1])
print('preprocess', d['ms_per_step'], {k: v['ms'] for k, v in d['kernels'].items() if k in ('conv1d_tc_kernel', 'shared_norm_planes_kernel')})
PY
