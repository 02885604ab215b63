#!/bin/bash
# ncu --set full of the secondary evaluators' kernels: usage gpu_ncu_evals.sh <tag>
TAG=${1:-ncue}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python profiles/tools/evaluators_time.py > $OUT/evaluators.txt 2> $OUT/err; cat $OUT/evaluators.txt | cut -c1-200
timeout 600 python profiles/tools/adaptive_time.py > $OUT/adaptive.txt 2>> $OUT/err; grep synthetic $OUT/adaptive.txt | cut -c1-250
for K in mpx_hess_kernel mpx_fgrad_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 8 -c 1 -f -o $OUT/prof_$K \
    python profiles/tools/evaluators_time.py > $OUT/ncu_$K.log 2>&1
done
for K in mpx_adapt_kernel mpx_adapt_hess_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 8 -c 1 -f -o $OUT/prof_$K \
    python profiles/tools/adaptive_time.py > $OUT/ncu_$K.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 378 -c 30 --csv --log-file $OUT/launches_adaptive.csv python profiles/tools/adaptive_time.py > /dev/null 2>&1
ls -la $OUT
