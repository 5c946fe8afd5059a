#!/bin/bash
# Round 2, GPU call AP (2 GPUs): orderly teardown (parallel.shutdown) in the DDP tests and the bench
out=gpurun_out/r2ap
mkdir -p $out
timeout 600 python -m pytest tests/test_train_ddp_gpu.py -x -q > $out/pytest_ddp.log 2>&1; echo "ddp tests rc=$?"
tail -3 $out/pytest_ddp.log
start=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 5 --warmup 3 --only train,fargan > $out/bench_n2.json 2> $out/bench_n2.err; echo "bench n=2 rc=$? in $(( $(date +%s) - start )) s"
grep -c "did not return" $out/bench_n2.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2ap/bench_n2.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['ms_per_step'])
for name, entry in d['secondary'].items():
    print(name, {k: entry.get(k) for k in ('value', 'unit', 'ms_per_step')})
PY
