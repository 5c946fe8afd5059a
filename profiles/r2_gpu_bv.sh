#!/bin/bash
# Round 2, GPU call BV: FARGAN without the local-memory pointer table and the generic loads: tests, same-box
# A/B against f74c0d8
out=gpurun_out/r2bv
mkdir -p $out
root=$PWD
timeout 900 python -m pytest tests/test_fargan_gpu.py -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest.log
for round in 1 2 3; do
  cd $root/profiles/debug/ab/before; timeout 300 python profiles/bench_fargan.py --steps 5 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('before', d['ms_per_step'])"
  cd $root; timeout 300 python profiles/bench_fargan.py --steps 5 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('head  ', d['ms_per_step'])"
done
