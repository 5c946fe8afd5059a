#!/bin/bash
# Round 2, GPU call A: tests, the bench line with every secondary workload, ncu captures of the
# non-convolution design choices, sanitizer passes.  Everything lands in gpurun_out/r2a/.
out=gpurun_out/r2a
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $out/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
tail -c 1500 $out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'viterbi_cluster_kernel|stft_kernel|pool_norm_planes_kernel|posterior_kernel|mel_kernel' \
    --launch-skip 0 -c 14 -o $out/preprocess -f python profiles/bench_preprocess.py --steps 1 --no-cpu > $out/ncu_pre.log 2>&1; echo "ncu pre rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fargan_kernel -c 1 \
    -o $out/fargan -f python profiles/bench_fargan.py --steps 1 --no-cpu > $out/ncu_fargan.log 2>&1; echo "ncu fargan rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file $out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $out/launches_bench.log 2>&1; echo "launch list rc=$?"
for file in test_conv1d_tc_gpu test_fargan_gpu; do
    timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/$file.py -x -q \
        > $out/memcheck_$file.log 2>&1; echo "memcheck $file rc=$?"
done
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_conv1d_tc_gpu.py -x -q -k "epilogue or golden" \
    > $out/racecheck_conv1d_tc.log 2>&1; echo "racecheck rc=$?"
ls -la $out
