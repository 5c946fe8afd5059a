#!/bin/bash
# Round 2, GPU call CJ (2 GPUs): data-parallel tests and the train line under torchrun after the training-kernel changes
out=gpurun_out/r2cj
mkdir -p $out
timeout 900 python -m pytest tests/test_train_ddp_gpu.py -q > $out/pytest_ddp.log 2>&1; echo "ddp pytest rc=$?"; tail -2 $out/pytest_ddp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2cj/bench_n2.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches')}, 'e2e', d['e2e']['value'])
for name, entry in d['secondary'].items():
    print(name, {k: entry.get(k) for k in ('value', 'unit', 'ms_per_step')})
PY
