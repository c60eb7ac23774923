#!/bin/bash
# Round profile capture (run under gpurun, one GPU): the launch list of a short bench run and one `ncu --set full`
# capture per hot kernel, each after the same command has exited 0 without ncu.  Reports land in gpurun_out/;
# tools/ncu_summary.py turns them into the summaries committed under profiles/.
#   usage: tools/profile_round.sh <tag>      e.g. r02a
set -u
TAG=${1:-r02}
OUT=gpurun_out
SHORT="--steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 20 --rollout-steps 16 --update-T 4 --update-dp 0 --state-warm 100 --other-configs 0"
python bench.py $SHORT > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain bench failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $SHORT > $OUT/${TAG}_ncu_l.log 2>&1
for K in fused_step_kernel policy_step_tc_kernel returns_kernel ppo_grad_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/${TAG}_$K python bench.py $SHORT > $OUT/${TAG}_ncu_$K.log 2>&1
done
SHORT5="--config cfg5 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 2 --rollout-steps 0 --update-T 0 --update-dp 0 --state-warm 20 --other-configs 0"
python bench.py $SHORT5 > $OUT/${TAG}_cfg5_plain.json 2> $OUT/${TAG}_cfg5_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:warp_step_kernel -s 3 -c 1 -f -o $OUT/${TAG}_warp_step_kernel python bench.py $SHORT5 > $OUT/${TAG}_ncu_warp.log 2>&1
ls -la $OUT/${TAG}_*
