#!/bin/bash
TAG=${1:-adapt3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file $OUT/adaptive_launches.csv python profiles/tools/adaptive_probe.py 3 > $OUT/ncu1.log 2>&1
grep -v "^==" $OUT/adaptive_launches.csv | cut -d, -f5,12-15 | tail -30
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mpx_adapt_kernel -s 1 -c 1 -f -o $OUT/prof_adapt python profiles/tools/adaptive_probe.py 3 > $OUT/ncu2.log 2>&1
ls -la $OUT
