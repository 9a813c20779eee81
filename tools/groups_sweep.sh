#!/bin/bash
# cfg4 batched throughput vs the number of concurrent sub-batches per GPU (device-resident and end to end)
for g in 2 4 8 16; do
  timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu --batch-groups $g 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['batched']
print('groups $g: device %.0f solves/s (%.1f ms), e2e %.0f solves/s (%.1f ms), converged %d' % (b['value'], 1e3*b['device_seconds'], b['e2e']['value'], 1e3*b['e2e']['seconds'], b['converged']))"
done
