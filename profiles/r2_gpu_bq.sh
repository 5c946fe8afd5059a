#!/bin/bash
# Round 2, GPU call BQ: pair kernel with pointer-increment addressing, new narrow-layer selection: tests,
# which blocks should now run fused, same-box A/B against 3bc6684
out=gpurun_out/r2bq
mkdir -p $out
root=$PWD
timeout 900 python -m pytest tests/test_conv1d_tc_gpu.py tests/test_generator_gpu.py tests/test_conv_pair_tc_gpu.py tests/test_benchmark_shapes_gpu.py -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
timeout 900 python profiles/pair_selection.py --steps 5 > $out/pair_selection.txt 2>&1; echo "pair selection rc=$?"; cat $out/pair_selection.txt
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
