#!/bin/bash
# Round 2, GPU call CK: which variant of the pair kernel is fastest now that its converters are cheaper?
mkdir -p gpurun_out/r2ck
timeout 300 python profiles/pair_variants.py > gpurun_out/r2ck/pair_variants.txt 2>&1; echo "rc=$?"; cat gpurun_out/r2ck/pair_variants.txt | grep -v Warning
