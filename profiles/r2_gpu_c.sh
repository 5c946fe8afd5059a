#!/bin/bash
# Round 2, GPU call C: pair-kernel variants (tests first), role breakdown, block selection
out=gpurun_out/r2c
mkdir -p $out
timeout 300 python -m pytest tests/test_conv_pair_tc_gpu.py -q > $out/pytest_pair.log 2>&1; rc=$?; echo "pair tests rc=$rc"
tail -15 $out/pytest_pair.log
timeout 900 python profiles/pair_breakdown.py > $out/pair_breakdown.txt 2>&1; echo "breakdown rc=$?"
cat $out/pair_breakdown.txt
timeout 600 python profiles/pair_selection.py > $out/pair_selection.txt 2>&1; echo "selection rc=$?"
cut -c1-330 $out/pair_selection.txt
ls -la $out
