#!/bin/bash
# Round 2, GPU call AN: compute-sanitizer over the kernels written in the second session
out=gpurun_out/r2an
mkdir -p $out
export PMN_TCW=2   # conv1d_tcw_kernel wherever it applies
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_conv1d_tc_gpu.py -x -q \
    > $out/memcheck_conv1d_tc_tcw.log 2>&1; echo "memcheck conv1d_tc (tcw everywhere) rc=$?"
tail -3 $out/memcheck_conv1d_tc_tcw.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_conv1d_tc_gpu.py -x -q -k "epilogue or golden or fp64" \
    > $out/racecheck_conv1d_tc_tcw.log 2>&1; echo "racecheck conv1d_tc (tcw everywhere) rc=$?"
tail -3 $out/racecheck_conv1d_tc_tcw.log
unset PMN_TCW
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_fargan_gpu.py -x -q \
    > $out/memcheck_fargan.log 2>&1; echo "memcheck fargan rc=$?"
tail -3 $out/memcheck_fargan.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_preprocess_gpu.py -x -q -k "viterbi or pitch_pipeline" \
    > $out/memcheck_preprocess.log 2>&1; echo "memcheck viterbi + pitch pipeline rc=$?"
tail -3 $out/memcheck_preprocess.log
