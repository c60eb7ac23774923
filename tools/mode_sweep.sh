#!/bin/bash
# eager vs device-round vs graph replay of the timed blocks (bench, rotating shards)
for m in "--graph 0" "--graph 0 --device-round" "--graph 1" "--graph 1 --streams 2" "--graph 1 --streams 1" "--graph 0 --streams 1"; do
  python bench.py --steps 1920 --warmup 20 --no-cpu-baseline --e2e-steps 0 --rollout-steps 0 $m "$@" 2>&1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m', d['kernels']['step_us'], d['roofline']['frac'])"
done
