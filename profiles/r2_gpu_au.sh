#!/bin/bash
# Round 2, GPU call AU: whole GPU suite again (pair-mask test fixed for the new kernel mix)
out=gpurun_out/r2au
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $out/pytest.log
