#!/bin/bash
# full GPU suite; the large-parity tests write gpurun_out/parity_large.json
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
