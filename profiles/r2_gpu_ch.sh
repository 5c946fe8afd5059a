#!/bin/bash
# Round 2, GPU call CH: wgrad split chosen by whole waves: training tests, train step time
out=gpurun_out/r2ch
mkdir -p $out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_train_ops_gpu.py -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest.log
for round in 1 2; do
timeout 600 python bench.py --only train --no-cpu-baseline > $out/bench_$round.json 2> $out/bench_$round.err; echo "rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('$out/bench_$round.json') if l.startswith('{')][-1])
t = d['secondary']['train']
print(round(t['ms_per_step'], 3), t['value'], {k: v['ms_per_step'] for k, v in sorted(t['roofline']['kernels'].items(), key=lambda x: -x[1]['ms_per_step'])[:9]})
PY
done
