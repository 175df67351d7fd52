#!/bin/bash
# One point of the 1 / 2 / 4 / 8-GPU tables (run on the GPU box: `gpurun --gpus N -- bash
# tools/run_scaling.sh N TAG`): bench.py (C2, weak + strong scaling, compact and full-record
# end to end) and every other BASELINE configuration through tools/bench_configs.py, whole job
# split over N ranks (C1 / C3 / C4 strided, C5 contiguous). One JSON line per run in
# gpurun_out/scale_<TAG>_n<N>.jsonl.
N=${1:-1}
TAG=${2:-r02}
OUT=gpurun_out/scale_${TAG}_n${N}.jsonl
ERR=gpurun_out/scale_${TAG}_n${N}.err
: > $OUT; : > $ERR
if [ "$N" -gt 1 ]; then
    RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
else
    RUN="python"
fi
$RUN bench.py --gpus $N --steps 3 --warmup 3 --cpu-rays $([ "$N" -gt 1 ] && echo 0 || echo 4194304) >> $OUT 2>> $ERR
BC="tools/bench_configs.py --no-cpu --steps 2"
$RUN $BC --config c1 --range 0 >> $OUT 2>> $ERR
$RUN $BC --config c1 --range 10 >> $OUT 2>> $ERR
$RUN $BC --config c3 --range 0 --rays 67108864 >> $OUT 2>> $ERR
$RUN $BC --config c3 --rays 67108864 >> $OUT 2>> $ERR
$RUN $BC --config c4 --range 0 >> $OUT 2>> $ERR
$RUN $BC --config c4 >> $OUT 2>> $ERR
$RUN $BC --config c4 --range 0 --multi 10 >> $OUT 2>> $ERR  # turtle_stepper_walk_batch
$RUN $BC --config c4 --multi 10 >> $OUT 2>> $ERR
$RUN $BC --config c5 >> $OUT 2>> $ERR
$RUN $BC --config c5 --gather 1 >> $OUT 2>> $ERR
cut -c1-260 $OUT
