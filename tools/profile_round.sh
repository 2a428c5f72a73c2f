#!/bin/bash
# Round profile artefacts (run on the GPU box through gpurun; outputs under gpurun_out/, summarised into profiles/ here
# with tools/ncu_summary.py).  usage: bash tools/profile_round.sh r02
set -u
TAG=${1:-r02}
OUT=gpurun_out
BENCH="python bench.py --steps 30 --warmup 5 --ff 128 --no-mcts --no-cpu-baseline --no-dropin"
$BENCH > $OUT/${TAG}_bench_short.log 2> $OUT/${TAG}_bench_short.err || { echo "bench failed"; exit 1; }
# (1) launch list of the same command: every launch with its device time (cold cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches_all.csv $BENCH > $OUT/${TAG}_ncu_launch.log 2>&1
# (2) the two kernels of a step, full sets (steps 3-4 of 6 after a 200-ply fast-forward)
python tools/ncu_step.py 200 > $OUT/${TAG}_step_plain.log 2>&1 || { echo "ncu_step failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'rules_kernel|expand_kernel' -s 204 -c 4 -f -o $OUT/${TAG}_step_full python tools/ncu_step.py 200 > $OUT/${TAG}_ncu_step.log 2>&1
# (3) the PUCT kernels, full sets (simulations 31-32 of 40: the trees are a few levels deep)
python tools/mcts_probe.py 40 > $OUT/${TAG}_mcts_plain.log 2>&1 || { echo "mcts_probe failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'tree_|rules_kernel|expand_kernel' -s 90 -c 6 -f -o $OUT/${TAG}_puct_full python tools/mcts_probe.py 40 > $OUT/${TAG}_ncu_puct.log 2>&1
tail -n 2 $OUT/${TAG}_step_plain.log $OUT/${TAG}_mcts_plain.log
