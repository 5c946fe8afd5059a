#!/bin/bash
# Round 2, GPU call BT: shared-memory accesses as LDS / STS instead of generic LD / ST in the three conv
# kernels: tests, narrow layers, same-box A/B against f74c0d8
out=gpurun_out/r2bt
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
for round in 1 2 3; do
  run $root/profiles/debug/ab/before before$round
  run $root head$round
done
timeout 600 python profiles/narrow_layers.py > $out/narrow_layers.txt 2>&1; echo "narrow rc=$?"; cat $out/narrow_layers.txt
timeout 900 python profiles/pair_selection.py --steps 5 > $out/pair_selection.txt 2>&1; echo "pair selection rc=$?"; python - <<'PY'
import json
for line in open('gpurun_out/r2bt/pair_selection.txt'):
    if line.startswith('{'):
        d = json.loads(line); print({k: d[k] for k in d if k in ('mask', 'channels', 'kernel', 'ms', 'gain_ms', 'best_mask')})
PY
