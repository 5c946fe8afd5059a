#!/bin/bash
# Round 2, GPU call AX: FARGAN, recurrent and side sums off the critical path
out=gpurun_out/r2ax
mkdir -p $out
timeout 300 python -m pytest tests/test_fargan_gpu.py -q -x --timeout 120 > $out/pytest.log 2>&1; echo "tests rc=$?"
tail -6 $out/pytest.log
PMN_FARGAN_DEBUG=1 timeout 300 python bench.py --steps 5 --warmup 3 --only fargan > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2ax/bench.json') if l.startswith('{')][-1])
f = d['secondary']['fargan']
print(f['value'], f['ms_per_step'], f.get('parity'), f['roofline'].get('us_per_subframe'))
PY
grep "fargan CTA 0" $out/bench.err | tail -1
