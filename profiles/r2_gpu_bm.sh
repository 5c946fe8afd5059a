#!/bin/bash
# Round 2, GPU call BM: same-box A/B of the epilogue pipeline (f3c9a45 = before, head = after), two
# alternating runs each, fp8 form on and off
out=gpurun_out/r2bm
mkdir -p $out
root=$PWD
run() {  # tree label f8
  cd $1
  PMN_GENERATOR_F8=$3 timeout 600 python bench.py --no-secondary --no-cpu-baseline > $root/$out/bench_$2_f8$3.json 2> $root/$out/bench_$2_f8$3.err
  cd $root
  python - <<PY
import json
d = json.loads([l for l in open('$out/bench_$2_f8$3.json') if l.startswith('{')][-1])
k = d['roofline']['kernels']
print('$2 f8=$3', round(d['ms_per_step'], 3), {n: round(k[n]['ms_per_step'], 3) for n in ('conv1d_tc_kernel', 'conv1d_tcw_kernel', 'conv_pair_tc_kernel')})
PY
}
for round in 1 2; do
  run $root/profiles/debug/ab/f3c9a45 before$round 1
  run $root head$round 1
  run $root/profiles/debug/ab/f3c9a45 before$round 0
  run $root head$round 0
done
timeout 600 python -m pytest tests/test_conv_pair_tc_gpu.py -q -x 2>&1 | tail -2
