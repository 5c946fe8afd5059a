#!/bin/bash
# Round 2, GPU call BY: end-of-round evidence at HEAD (f74c0d8+): sanitizers over the generator and conv
# tests, the whole GPU suite, the default bench line, the reference arm, smoke, launch list of one forward,
# ncu --set full of one preprocess step, ncu --set full of the shipped FARGAN kernel and block-1 operand writer
out=gpurun_out/r2by
mkdir -p $out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_generator_gpu.py -x -q \
    -k "matches_oracle or golden" > $out/memcheck_generator.log 2>&1; echo "memcheck generator rc=$?"
tail -2 $out/memcheck_generator.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_conv1d_tc_gpu.py tests/test_conv_pair_tc_gpu.py -x -q \
    -k "not 38403 and not 35847 and not many_tiles" > $out/memcheck_conv.log 2>&1; echo "memcheck conv1d_tc + pair rc=$?"
tail -2 $out/memcheck_conv.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_conv1d_tc_gpu.py tests/test_conv_pair_tc_gpu.py -x -q \
    -k "epilogue or golden or accumulate" > $out/racecheck_conv.log 2>&1; echo "racecheck conv1d_tc + pair rc=$?"
tail -2 $out/racecheck_conv.log
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $out/pytest.log
start=$(date +%s)
timeout 1200 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$? in $(( $(date +%s) - start )) s"
timeout 600 python bench.py --impl reference > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -c 400 --csv --log-file $out/forward_dram_traffic.csv python profiles/forward_once.py > $out/forward_once.log 2>&1; echo "traffic rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file $out/preprocess_launches.csv python profiles/bench_preprocess.py --steps 1 --no-cpu > $out/preprocess_once.log 2>&1; echo "preprocess launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fargan_kernel -c 1 \
    -o $out/fargan -f python profiles/bench_fargan.py --steps 1 --no-cpu > $out/ncu_fargan.log 2>&1; echo "ncu fargan rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:shared_norm_planes_kernel -c 1 \
    -o $out/shared_norm -f python profiles/bench_preprocess.py --steps 1 --no-cpu > $out/ncu_norm.log 2>&1; echo "ncu shared_norm rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2by/bench.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('metric', 'value', 'ms_per_step', 'steps', 'warmup', 'gpu_launches', 'clocks')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['traffic'], 'parity', d['parity']['max_rel_err'])
for k, v in sorted(d['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:6]: print('  ', k, v)
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
for name, entry in d['secondary'].items():
    print(name, {k: entry.get(k) for k in ('value', 'unit', 'ms_per_step')}, entry.get('parity'))
r = json.loads([l for l in open('gpurun_out/r2by/bench_ref.json') if l.startswith('{')][-1])
print('reference', {k: r.get(k) for k in ('value', 'ms_per_step', 'steps')})
PY
