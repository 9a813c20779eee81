#!/bin/bash
mkdir -p gpurun_out


timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02a.log 2> gpurun_out/bench_r02a.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r02a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02a.log').read().strip().splitlines()[-1])
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
print("batched", json.dumps(d["batched"], indent=1)[:3000])
print("configs", json.dumps(d["configs"], indent=1)[:2500])
print("cpu", d["cpu_baseline"])
PY
