#!/bin/bash
# One full ncu capture of the g+jac kernel (2 launches after warm-up): usage gpu_ncu.sh <tag> [env assignments...]
TAG=${1:-ncu}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k regex:mpx_gjac -s 3 -c 2 -f -o $OUT/prof_gjac \
    python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log | cut -c1-300
