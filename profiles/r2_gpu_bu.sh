#!/bin/bash
# Round 2, GPU call BU: default pair mask 0x248 against 0x008, three alternating bench runs; tests that
# depend on the mask
out=gpurun_out/r2bu
mkdir -p $out
timeout 900 python -m pytest tests/test_generator_gpu.py tests/test_conv_pair_tc_gpu.py tests/test_benchmark_shapes_gpu.py -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
for round in 1 2 3; do
for mask in 0x008 0x248; do
  PMN_PAIR_MASK=$mask timeout 600 python bench.py --no-secondary --no-cpu-baseline > $out/bench_${mask}_$round.json 2> $out/bench_${mask}_$round.err
  python - <<PY
import json
d = json.loads([l for l in open('$out/bench_${mask}_$round.json') if l.startswith('{')][-1])
k = d['roofline']['kernels']
print('$mask run $round', round(d['ms_per_step'], 3), 'parity', d['parity']['max_rel_err'], {n: round(k[n]['ms_per_step'], 3) for n in ('conv1d_tc_kernel', 'conv1d_tcw_kernel', 'conv_pair_tc_kernel')})
PY
done
done
