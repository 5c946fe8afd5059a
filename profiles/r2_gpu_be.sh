#!/bin/bash
# Round 2, GPU call BE: 16 epilogue warps for the C = 32 launches of conv1d_tc_kernel
out=gpurun_out/r2be
mkdir -p $out
timeout 900 python -m pytest tests/test_generator_gpu.py tests/test_conv1d_tc_gpu.py tests/test_conv_pair_tc_gpu.py tests/test_preprocess_gpu.py -q --timeout 300 > $out/pytest.log 2>&1; echo "tests rc=$?"
tail -3 $out/pytest.log
timeout 900 python profiles/narrow_layers.py | head -14 | tee $out/narrow_layers.txt
timeout 600 python bench.py --steps 10 --warmup 3 --only synthesis > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2be/bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['parity']['max_rel_err'])
PY
