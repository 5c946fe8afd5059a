#!/bin/bash
# Round 2, GPU call CG: launch list of one eager training step, then ncu --set full of the heaviest forward /
# data-gradient and weight-gradient launches
out=gpurun_out/r2cg
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 900 --csv \
    --log-file $out/train_launches.csv python profiles/train_step_once.py > $out/once.log 2>&1; echo "launch list rc=$?"
python - <<'PY' > gpurun_out/r2cg/heaviest.txt
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2cg/train_launches.csv')) if len(r) > 14 and r[0].isdigit()]
best = {}
for r in rows:
    name = 'gemm' if 'conv_gemm_tc_kernel' in r[4] else 'wgrad' if 'conv_wgrad_tc_kernel' in r[4] else None
    if name and (name not in best or float(r[14]) > best[name][1]):
        best[name] = (int(r[0]), float(r[14]), r[4][:90], r[8])
for name, value in best.items():
    print(name, *value)
PY
cat $out/heaviest.txt
gemm=$(grep '^gemm' $out/heaviest.txt | awk '{print $2}')
wgrad=$(grep '^wgrad' $out/heaviest.txt | awk '{print $2}')
timeout 600 ncu --set full --clock-control none --import-source on -s $gemm -c 1 -o $out/gemm -f python profiles/train_step_once.py > $out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -s $wgrad -c 1 -o $out/wgrad -f python profiles/train_step_once.py > $out/ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
