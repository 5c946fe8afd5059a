#!/bin/bash
# Round 2, GPU call BX: pitch block 1 in the fp8 form with the operand writer at two CTAs per SM
out=gpurun_out/r2bx
mkdir -p $out
timeout 600 python -m pytest tests/test_preprocess_gpu.py -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest.log
for round in 1 2; do
for f8 in 0 1; do
PMN_PITCH_F8=$f8 timeout 600 python profiles/bench_preprocess.py --steps 5 --no-cpu > $out/preprocess_f8_${f8}_$round.json 2> $out/preprocess_f8_${f8}_$round.err
python - <<PY
import json
d = json.loads([l for l in open('$out/preprocess_f8_${f8}_$round.json') if l.startswith('{')][-1])
print('pitch f8=$f8', round(d['ms_per_step'], 3), {k: v['ms'] for k, v in d['kernels'].items() if k in ('conv1d_tc_kernel', 'shared_norm_planes_kernel')})
PY
done
done
