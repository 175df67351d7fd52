#!/bin/bash
# C4 through turtle_stepper_walk_batch (run on the GPU box): steps per launch 1 / 10 / 100 at
# both ranges, then the whole walk in one launch with the reference on the host cores and
# parity. One JSON line per run in gpurun_out/walk_multi_<TAG>.jsonl.
TAG=${1:-r02}
MULTIS=${2:-"1 10 100"}
OUT=gpurun_out/walk_multi_$TAG.jsonl
ERR=gpurun_out/walk_multi_$TAG.err
: > $OUT; : > $ERR
for rg in 0 1; do
    for m in $MULTIS; do
        timeout 150 python tools/bench_configs.py --config c4 --range $rg --multi $m --no-cpu >> $OUT 2>> $ERR
    done
done
timeout 200 python tools/bench_configs.py --config c4 --multi 100 >> $OUT 2>> $ERR
timeout 200 python tools/bench_configs.py --config c4 --range 0 --multi 100 >> $OUT 2>> $ERR
cut -c1-330 $OUT
