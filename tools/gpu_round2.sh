#!/bin/bash
# Round-2 evidence pass on one B200: GPU parity suite, bench line, reference arm, phase clocks of the fused batch
# kernel, launch lists (bench step + one batched solve), ncu --set full captures of k_ipm_batch and k_ldlt_dataflow.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -16 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
for c in 4096 512; do
  echo "== clk build $c"; IPMZ_LIB=$PWD/ipm-zoo_b200/ab/libipmz_clk.so timeout 300 python tools/prof_batched.py $c 2>&1 | tail -4
  echo "== product $c"; timeout 300 python tools/prof_batched.py $c 2>&1 | tail -2
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-batched --no-configs > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_batched_4096.csv python tools/prof_batched.py 4096 > gpurun_out/ncu_launch_b.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_ipm_batch -s 1 -c 1 -o gpurun_out/prof_batch -f python tools/prof_batched.py 592 > gpurun_out/ncu_batch.log 2>&1
tail -2 gpurun_out/ncu_batch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ldlt_dataflow|k_trsv_fused" -s 2 -c 2 -o gpurun_out/prof_dataflow -f python tools/prof_factor.py 8192 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
