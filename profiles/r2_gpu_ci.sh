#!/bin/bash
# Round 2, GPU call CI: the whole GPU suite, default bench line, reference arm and smoke at the final HEAD
# (after the training-kernel changes); memcheck over the training-op tests
out=gpurun_out/r2ci
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $out/pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_train_ops_gpu.py -x -q \
    > $out/memcheck_train_ops.log 2>&1; echo "memcheck train ops rc=$?"; tail -2 $out/memcheck_train_ops.log
start=$(date +%s)
timeout 1200 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$? in $(( $(date +%s) - start )) s"
timeout 600 python bench.py --impl reference > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2ci/bench.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('metric', 'value', 'ms_per_step', 'steps', 'warmup', 'gpu_launches', 'clocks')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['traffic'], 'parity', d['parity']['max_rel_err'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
for name, entry in d['secondary'].items():
    print(name, {k: entry.get(k) for k in ('value', 'unit', 'ms_per_step')}, entry.get('parity'))
print('eager', {k: v.get('ms_per_step') for k, v in d['gpu_eager_baseline']['modes'].items()}, {k: v.get('ms_per_step') for k, v in d['gpu_eager_baseline']['train']['modes'].items()}, d['gpu_eager_baseline'].get('fargan', {}).get('ms_per_step'))
r = json.loads([l for l in open('gpurun_out/r2ci/bench_ref.json') if l.startswith('{')][-1])
print('reference', {k: r.get(k) for k in ('value', 'ms_per_step', 'steps')})
PY
