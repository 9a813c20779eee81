#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
tail -c 400 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${N}gpu.log') if l.startswith('{')][-1])
print("N", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"])
b=d["batched"]; print("batched dev", b["value"], "e2e", b["e2e"]["value"], "floor", b["e2e"]["floor_seconds"], "e2e_s", b["e2e"]["seconds"], "up_s", b["e2e"]["upload_only_seconds"], "h2d/rank", b["e2e"]["h2d_gbs_per_rank"], "ratio", b["e2e"]["e2e_over_floor"], "dev_s", b["device_seconds"])
PY
