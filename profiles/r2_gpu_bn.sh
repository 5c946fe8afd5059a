#!/bin/bash
# Round 2, GPU call BN: compute-sanitizer over the reworked conv1d_tc_kernel (three warpgroups, setmaxnreg,
# residual pipeline across tiles, fp8 operand form), then the whole GPU suite, bench, reference arm, smoke
out=gpurun_out/r2bn
mkdir -p $out
export PMN_TCW=0   # every shape on conv1d_tc_kernel
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_conv1d_tc_gpu.py -x -q \
    -k "not 38403 and not 35847" > $out/memcheck_conv1d_tc.log 2>&1; echo "memcheck conv1d_tc rc=$?"
tail -3 $out/memcheck_conv1d_tc.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_conv1d_tc_gpu.py -x -q \
    -k "epilogue or golden or (fp64 and not 38403 and not 35847 and not 2000)" > $out/racecheck_conv1d_tc.log 2>&1; echo "racecheck conv1d_tc rc=$?"
tail -3 $out/racecheck_conv1d_tc.log
unset PMN_TCW
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_generator_gpu.py -x -q \
    -k "matches_oracle or golden" > $out/memcheck_generator.log 2>&1; echo "memcheck generator rc=$?"
tail -3 $out/memcheck_generator.log
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $out/pytest.log
start=$(date +%s)
timeout 1200 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$? in $(( $(date +%s) - start )) s"
timeout 600 python bench.py --impl reference > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2bn/bench.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('metric', 'value', 'ms_per_step', 'steps', 'warmup', 'gpu_launches', 'clocks', 'dtype')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['traffic'], 'parity', d['parity']['max_rel_err'])
for k, v in sorted(d['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:6]: print('  ', k, v)
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
for name, entry in d['secondary'].items():
    print(name, {k: entry.get(k) for k in ('value', 'unit', 'ms_per_step')}, entry.get('parity'))
print('eager', {k: v.get('ms_per_step') for k, v in d['gpu_eager_baseline']['modes'].items()}, {k: v.get('ms_per_step') for k, v in d['gpu_eager_baseline']['train']['modes'].items()}, d['gpu_eager_baseline'].get('fargan', {}).get('ms_per_step'))
r = json.loads([l for l in open('gpurun_out/r2bn/bench_ref.json') if l.startswith('{')][-1])
print('reference', {k: r.get(k) for k in ('value', 'ms_per_step', 'steps')})
PY
