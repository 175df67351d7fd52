#!/bin/bash
# ncu --set full captures of every kernel that carries a BASELINE configuration (run on the
# GPU box: `gpurun -- bash tools/capture_profiles.sh TAG`). One launch each, after a plain run
# of the same command has exited 0; the .ncu-rep files come back in gpurun_out/ and are
# turned into profiles/r02_kernels.json + profiles/r02_<name>.md by tools/kernel_profiles.py.
TAG=${1:-r02}
ONLY=${2:-all} # a comma-separated subset of the names below (the .ncu-rep files of one
               # gpurun call must stay under 64 MiB: three or four captures per call)
BC="python tools/bench_configs.py"
NCU="ncu --set full --import-source on --clock-control none --launch-count 1 -f"
O=gpurun_out
run() { # name, kernel regex, launches to skip, command...
    name=$1; regex=$2; skip=$3; shift 3
    case ",$ONLY," in *",all,"*|*",$name,"*) ;; *) return;; esac
    if [ "$(du -sm $O | cut -f1)" -gt 48 ]; then echo "$name: skipped, $O is near the 64 MiB that come back"; return; fi
    "$@" > $O/plain_${name}_$TAG.json 2> $O/plain_${name}_$TAG.err || { echo "$name: plain run failed"; return; }
    $NCU -k regex:$regex --launch-skip $skip -o $O/prof_${name}_$TAG "$@" > $O/ncu_${name}_$TAG.log 2>&1
    tail -1 $O/plain_${name}_$TAG.json | cut -c1-300
}
run c2_trace_stack trace_kernel 3 python bench.py --steps 1 --no-e2e --cpu-rays 0
run c1_trace_proj trace_kernel 3 $BC --config c1 --range 0 --no-cpu --steps 1
run c1_trace_lla_proj trace_kernel 3 $BC --config c1 --range 10 --no-cpu --steps 1
run c3_trace_proj trace_kernel 3 $BC --config c3 --range 0 --rays 2097152 --max-steps 3000 --no-cpu --steps 1
run c3_trace_lla_proj trace_kernel 3 $BC --config c3 --rays 2097152 --max-steps 3000 --no-cpu --steps 1
run c4_walk_lla_proj walk_kernel 9 $BC --config c4 --rays 4194304 --walk 6 --no-cpu
run c4_walk_proj walk_kernel 9 $BC --config c4 --range 0 --rays 4194304 --walk 6 --no-cpu
run c5_to_geodetic to_geodetic_kernel 1 $BC --config c5 --rays 268435456 --steps 1 --warmup 1 --no-cpu
run c5_map_elevation "map_elevation_kernel" 1 $BC --config c5 --rays 268435456 --steps 1 --warmup 1 --no-cpu
run c5_map_elevation_ecef map_elevation_ecef_kernel 1 $BC --config c5 --rays 268435456 --steps 1 --warmup 1 --no-cpu
run c5_map_elevation_packed "map_elevation_kernel" 1 $BC --config c5 --gather 1 --rays 268435456 --steps 1 --warmup 1 --no-cpu
run c5_map_elevation_ecef_packed map_elevation_ecef_kernel 1 $BC --config c5 --gather 1 --rays 268435456 --steps 1 --warmup 1 --no-cpu
run c4_walk_multi_proj walk_kernel 3 $BC --config c4 --range 0 --rays 4194304 --walk 12 --multi 6 --no-cpu
run c4_walk_multi_lla_proj walk_kernel 3 $BC --config c4 --rays 4194304 --walk 12 --multi 6 --no-cpu
