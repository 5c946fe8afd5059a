#!/bin/bash
# Round 2, GPU call Y: the whole GPU suite, the default bench line and the reference arm
out=gpurun_out/r2y
mkdir -p $out
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $out/pytest.log
fi
start=$(date +%s)
timeout 1200 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$? in $(( $(date +%s) - start )) s"
tail -c 600 $out/bench.err
timeout 600 python bench.py --impl reference > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2y/bench.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('metric', 'value', 'ms_per_step', 'steps', 'warmup', 'gpu_launches', 'clocks')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'], 'parity', d['parity'])
print('cpu', d['cpu_baseline'])
for name, entry in d['secondary'].items():
    print(name, {k: entry.get(k) for k in ('metric', 'value', 'unit', 'ms_per_step')}, entry.get('parity'))
r = json.loads([l for l in open('gpurun_out/r2y/bench_ref.json') if l.startswith('{')][-1])
print('reference', {k: r.get(k) for k in ('value', 'ms_per_step', 'steps', 'cpu_baseline')})
PY
