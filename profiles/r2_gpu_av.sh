#!/bin/bash
# Round 2, GPU call AV: compute-sanitizer over the training kernels' unit tests
out=gpurun_out/r2av
mkdir -p $out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_train_ops_gpu.py -x -q \
    > $out/memcheck_train_ops.log 2>&1; echo "memcheck train ops rc=$?"
tail -3 $out/memcheck_train_ops.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_train_ops_gpu.py -x -q -k "exact or integers" \
    > $out/racecheck_train_ops.log 2>&1; echo "racecheck train ops rc=$?"
tail -3 $out/racecheck_train_ops.log
