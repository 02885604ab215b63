#!/bin/bash
TAG=${1:-adapt2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_adaptive.py -q -m gpu -x ) > $OUT/pytest_adapt.log 2>&1
tail -15 $OUT/pytest_adapt.log | cut -c1-300
timeout 300 python profiles/tools/adaptive_time.py > $OUT/adaptive_time.txt 2> $OUT/adaptive_time.err; cat $OUT/adaptive_time.txt; tail -5 $OUT/adaptive_time.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 12 --csv --log-file $OUT/adaptive_launches.csv python profiles/tools/adaptive_probe.py 1 > $OUT/ncu1.log 2>&1
