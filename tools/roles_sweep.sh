#!/bin/bash
# Throughput of the fused step kernel for 1/2/4 role warps per tile (bench, rotating shards).
for r in 1 2 4; do
  MSCHED_ROLES=$r python bench.py --steps 1920 --warmup 20 --no-cpu-baseline --e2e-steps 0 --rollout-steps 0 "$@" 2>&1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('roles $r', d['kernels']['step_us'], d['value'], d['roofline']['frac'])"
done
