#!/bin/bash
# Round 2, GPU call AC: conv1d_tcw_kernel with four epilogue warp sets
out=gpurun_out/r2ac
mkdir -p $out
timeout 900 python -m pytest tests/test_conv1d_tc_gpu.py tests/test_generator_gpu.py -q -x --timeout 300 > $out/pytest.log 2>&1; echo "tests rc=$?"
tail -3 $out/pytest.log
timeout 900 python profiles/narrow_layers.py | tee $out/narrow_layers.txt
