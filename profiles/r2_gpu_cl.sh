#!/bin/bash
# Round 2, GPU call CL: memcheck over the pitch-network tests (fp8-form operand writer, conv1d_tc frame mode) and FARGAN
out=gpurun_out/r2cl
mkdir -p $out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_preprocess_gpu.py -x -q -k "pitch or penn" \
    > $out/memcheck_pitch.log 2>&1; echo "memcheck pitch rc=$?"; tail -2 $out/memcheck_pitch.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_fargan_gpu.py -x -q \
    > $out/memcheck_fargan.log 2>&1; echo "memcheck fargan rc=$?"; tail -2 $out/memcheck_fargan.log
