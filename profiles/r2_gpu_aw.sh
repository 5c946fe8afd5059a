#!/bin/bash
# Round 2, GPU call AW: more shapes for the narrow layers' kernel, as shipped and with it forced everywhere
out=gpurun_out/r2aw
mkdir -p $out
timeout 600 python -m pytest tests/test_conv1d_tc_gpu.py tests/test_conv_pair_tc_gpu.py -q > $out/pytest.log 2>&1; echo "default rc=$?"
tail -3 $out/pytest.log
PMN_TCW=2 timeout 600 python -m pytest tests/test_conv1d_tc_gpu.py -q > $out/pytest_tcw.log 2>&1; echo "tcw everywhere rc=$?"
tail -3 $out/pytest_tcw.log
PMN_TCW=0 timeout 600 python -m pytest tests/test_conv1d_tc_gpu.py tests/test_conv_pair_tc_gpu.py -q > $out/pytest_tc.log 2>&1; echo "tcw off rc=$?"
tail -3 $out/pytest_tc.log
