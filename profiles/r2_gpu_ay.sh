#!/bin/bash
# Round 2, GPU call AY: final check of the round: whole GPU suite, default bench line, reference arm, smoke
out=gpurun_out/r2ay
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $out/pytest.log
start=$(date +%s)
timeout 1200 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$? in $(( $(date +%s) - start )) s"
timeout 600 python bench.py --impl reference > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2ay/bench.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('metric', 'value', 'ms_per_step', 'steps', 'warmup', 'gpu_launches', 'clocks')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['traffic'], 'parity', d['parity']['max_rel_err'])
for k, v in sorted(d['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:6]: print('  ', k, v)
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
for name, entry in d['secondary'].items():
    print(name, {k: entry.get(k) for k in ('value', 'unit', 'ms_per_step')}, entry.get('parity'))
print('eager', {k: v.get('ms_per_step') for k, v in d['gpu_eager_baseline']['modes'].items()}, {k: v.get('ms_per_step') for k, v in d['gpu_eager_baseline']['train']['modes'].items()}, d['gpu_eager_baseline'].get('fargan'))
r = json.loads([l for l in open('gpurun_out/r2ay/bench_ref.json') if l.startswith('{')][-1])
print('reference', {k: r.get(k) for k in ('value', 'ms_per_step', 'steps')})
PY
