#!/bin/bash
# Round 2, GPU call AD: conv1d_tcw_kernel, 4 epilogue sets at C = 32 and 2 at C = 64; selection by shape
out=gpurun_out/r2ad
mkdir -p $out
PMN_TCW=2 timeout 900 python -m pytest tests/test_conv1d_tc_gpu.py tests/test_generator_gpu.py -q -x --timeout 300 > $out/pytest_all.log 2>&1; echo "tests (tcw wherever it applies) rc=$?"
tail -2 $out/pytest_all.log
timeout 900 python -m pytest tests/test_conv1d_tc_gpu.py tests/test_generator_gpu.py tests/test_benchmark_shapes_gpu.py -q -x --timeout 300 -k "not train and not preprocess" > $out/pytest.log 2>&1; echo "tests (default) rc=$?"
tail -2 $out/pytest.log
timeout 900 python profiles/narrow_layers.py | tee $out/narrow_layers.txt
for flag in 0 1; do
PMN_TCW=$flag timeout 600 python bench.py --steps 10 --warmup 3 --only synthesis > $out/bench_tcw$flag.json 2> $out/bench_tcw$flag.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/bench_tcw$flag.json') if l.startswith('{')][-1])
print('PMN_TCW=$flag', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['parity']['max_rel_err'])
PY
done
