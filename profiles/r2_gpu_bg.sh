#!/bin/bash
# Round 2, GPU call BG: "fp16 + 2 x fp8" operands in ONE accumulator (scaled main term): conv tests,
# generator at the benchmark shape with three seeds, timing of synthesis and preprocess with it
out=gpurun_out/r2bg
mkdir -p $out
timeout 600 python -m pytest tests/test_conv1d_tc_gpu.py -q -x -k "f8" > $out/pytest_conv.log 2>&1; echo "conv f8 rc=$?"; tail -5 $out/pytest_conv.log
timeout 600 python -m pytest tests/test_preprocess_gpu.py -q -x -k "fp8" > $out/pytest_pitch.log 2>&1; echo "pitch f8 rc=$?"; tail -3 $out/pytest_pitch.log
timeout 900 python -m pytest tests/test_benchmark_shapes_gpu.py -q -x -s -k "fp8" > $out/pytest_gen.log 2>&1; echo "generator f8 rc=$?"; grep -E "relative errors|passed|failed" $out/pytest_gen.log
for f8 in 0 1; do
PMN_GENERATOR_F8=$f8 timeout 600 python bench.py --no-secondary --no-cpu-baseline > $out/bench_f8_$f8.json 2> $out/bench_f8_$f8.err; echo "bench f8=$f8 rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/bench_f8_$f8.json') if l.startswith('{')][-1])
print('f8=$f8', d['ms_per_step'], d['value'], 'parity', d['parity']['max_rel_err'], 'frac', d['roofline']['frac'])
for k, v in sorted(d['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:4]: print('  ', k, v)
PY
done
for f8 in 0 1; do
PMN_PITCH_F8=$f8 timeout 600 python profiles/bench_preprocess.py --steps 5 --no-cpu > $out/preprocess_f8_$f8.json 2> $out/preprocess_f8_$f8.err; echo "preprocess f8=$f8 rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/preprocess_f8_$f8.json') if l.startswith('{')][-1])
print('pitch f8=$f8', d['ms_per_step'], {k: v['ms'] for k, v in d['kernels'].items() if k in ('conv1d_tc_kernel', 'shared_norm_planes_kernel')})
PY
done
