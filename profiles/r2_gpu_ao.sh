#!/bin/bash
# Round 2, GPU call AO (2 GPUs): destroy_process_group with and without releasing the peer mappings
out=gpurun_out/r2ao
mkdir -p $out
for mode in release keep; do
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    profiles/debug/destroy_probe.py $mode > $out/destroy_$mode.log 2>&1; echo "$mode rc=$?"
grep -E "rank [01]:" $out/destroy_$mode.log
done
