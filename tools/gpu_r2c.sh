#!/bin/bash
# full GPU suite + bench line + launch list of one batched solve + ncu capture of k_ipm_batch
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full.log').read().strip().splitlines()[-1])
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["seconds"])
b=d["batched"]; print("batched", b["value"], "e2e", b["e2e"]["value"], "floor", b["e2e"]["floor_seconds"], b["e2e"]["seconds"], "gate", b.get("parity_gate"))
print("configs", {k:(v["e2e_ms"], v["iterations"]) for k,v in d["configs"].items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_batched_4096.csv python tools/prof_batched.py 4096 > gpurun_out/ncu_launch_b.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_ipm_batch -s 1 -c 1 -o gpurun_out/prof_batch -f python tools/prof_batched.py 592 > gpurun_out/ncu_batch.log 2>&1
tail -2 gpurun_out/ncu_batch.log
