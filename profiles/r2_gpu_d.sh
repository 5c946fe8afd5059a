#!/bin/bash
# Round 2, GPU call D (2 GPUs): the bench line under torchrun (every secondary workload at N = 2,
# training with the peer-memory exchange), and the 2-GPU tests the 1-GPU box skips
out=gpurun_out/r2d
mkdir -p $out
nvidia-smi -L > $out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err; echo "bench n2 rc=$?"
tail -c 600 $out/bench_n2.err
cut -c1-300 $out/bench_n2.json
timeout 600 python -m pytest tests/test_train_ddp_gpu.py -q > $out/pytest_2gpu.log 2>&1; echo "2-gpu tests rc=$?"
tail -8 $out/pytest_2gpu.log
ls -la $out
