#!/bin/bash
# Round 2, GPU call B: the fused pair kernel (tests under a short timeout: a wrong barrier phase
# would hang), block selection, ncu captures of the pair kernel and of the Viterbi kernel
out=gpurun_out/r2b
mkdir -p $out
timeout 300 python -m pytest tests/test_conv_pair_tc_gpu.py -q > $out/pytest_pair.log 2>&1; rc=$?; echo "pair tests rc=$rc"
tail -25 $out/pytest_pair.log
if [ $rc -ne 0 ]; then
    timeout 120 python -m pytest tests/test_conv_pair_tc_gpu.py -q -k "fp64" > $out/pytest_pair_all.log 2>&1
    tail -40 $out/pytest_pair_all.log
fi
if [ $rc -eq 0 ]; then
    timeout 600 python profiles/pair_selection.py > $out/pair_selection.txt 2>&1; echo "selection rc=$?"
    cat $out/pair_selection.txt
    timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
    cut -c1-700 $out/bench.json
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_pair_tc_kernel' -c 27 \
        -o $out/pair -f python profiles/forward_once.py > $out/ncu_pair.log 2>&1; echo "ncu pair rc=$?"
    timeout 600 python -m pytest tests/test_generator_gpu.py tests/test_benchmark_shapes_gpu.py -x -q -k "not preprocess and not train" > $out/pytest_generator.log 2>&1; echo "generator tests rc=$?"
    tail -5 $out/pytest_generator.log
fi
ls -la $out
