#!/bin/bash
# ncu --set full of one kernel of a tool script: usage gpu_ncu_one.sh <tag> <kernel regex> <script> [skip]
TAG=$1; K=$2; SCRIPT=$3; SKIP=${4:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o $OUT/prof_$K python $SCRIPT > $OUT/ncu_$K.log 2>&1
tail -2 $OUT/ncu_$K.log | cut -c1-200
