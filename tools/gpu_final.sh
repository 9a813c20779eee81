mkdir -p gpurun_out
echo "== create timing, pooled"; python tools/create_timing.py 2>&1 | tail -3
echo "== create timing, cudaMalloc"; IPMZ_POOL_ALLOC=0 python tools/create_timing.py 2>&1 | tail -3
for i in 1 2 3; do echo "fifo"; python tools/prof_batched.py 4096 | tail -1; echo "tickets"; IPMZ_FUSED_QUEUE=0 python tools/prof_batched.py 4096 | tail -1; done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_batched_4096.csv python tools/prof_batched.py 4096 > gpurun_out/ncu_launch_b.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-batched --no-configs > gpurun_out/ncu_launch.log 2>&1
