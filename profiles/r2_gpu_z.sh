#!/bin/bash
# Round 2, GPU call Z (8 GPUs): the driver's own multi-GPU launch of bench.py at N = 8 and N = 4
out=gpurun_out/r2z
mkdir -p $out
nvidia-smi -L | head -8
for n in 8 4; do
start=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 10 --warmup 3 > $out/bench_n$n.json 2> $out/bench_n$n.err; echo "bench n=$n rc=$? in $(( $(date +%s) - start )) s"
python - <<PY
import json
lines = [l for l in open('$out/bench_n$n.json') if l.startswith('{')]
print(len(lines), 'json lines')
d = json.loads(lines[-1])
print({k: d[k] for k in ('n_gpus', 'value', 'ms_per_step')}, 'e2e', d['e2e']['value'])
for name, entry in d['secondary'].items():
    print(name, {k: entry.get(k) for k in ('value', 'unit', 'ms_per_step', 'exchange')})
PY
done
start=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $out/bench_ref_n8.json 2> $out/bench_ref_n8.err; echo "ref n=8 rc=$? in $(( $(date +%s) - start )) s"
tail -c 300 $out/bench_ref_n8.json
timeout 900 python -m pytest tests/test_train_ddp_gpu.py -x -q > $out/pytest_ddp.log 2>&1; echo "ddp tests rc=$?"
tail -3 $out/pytest_ddp.log
