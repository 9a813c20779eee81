#!/bin/bash
# Round-2 evidence pass on one B200: GPU parity suite, bench line, reference arm, phase clocks of the fused batch
# kernel, launch lists (bench step + one batched solve) and one ncu --set full capture of k_ipm_batch.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/bench_full.log
tail -c 600 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
tail -c 1500 gpurun_out/bench_ref.log
for c in 4096 512; do
  echo "== clk build $c"; IPMZ_LIB=$PWD/ipm-zoo_b200/ab/libipmz_clk.so timeout 300 python tools/prof_batched.py $c 2>&1 | tail -4
  echo "== product $c"; timeout 300 python tools/prof_batched.py $c 2>&1 | tail -2
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-batched --no-configs > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_batched_4096.csv python tools/prof_batched.py 4096 > gpurun_out/ncu_launch_b.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_ipm_batch -s 1 -c 1 -o gpurun_out/prof_batch -f python tools/prof_batched.py 592 > gpurun_out/ncu_batch.log 2>&1
tail -3 gpurun_out/ncu_batch.log
ls -la gpurun_out
