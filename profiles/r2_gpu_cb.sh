#!/bin/bash
# Round 2, GPU call CB: weight-norm backward of a whole module in one launch: training tests, train step time
out=gpurun_out/r2cb
mkdir -p $out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_train_ops_gpu.py tests/test_benchmark_shapes_gpu.py -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest.log
for round in 1 2; do
timeout 600 python profiles/bench_train.py 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('ms_per_step','value','steps')}, {k: v for k, v in list(d.get('kernels', {}).items())[:6]})"
done
