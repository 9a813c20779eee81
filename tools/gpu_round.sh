#!/bin/bash
# One GPU-box pass: parity tests, the bench line, the ncu launch list and one full capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$?"
tail -c 3500 gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-batched > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ldlt_dataflow|k_trsv_fused" -s 2 -c 2 -o gpurun_out/prof_dataflow_r01 -f python tools/prof_factor.py 8192 2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bk_factor" -c 1 -o gpurun_out/prof_bk_r01 -f python tools/bk_time.py 1536 > gpurun_out/ncu_bk.log 2>&1
tail -3 gpurun_out/ncu_bk.log
