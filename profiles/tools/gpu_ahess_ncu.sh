#!/bin/bash
# one full ncu capture of mpx_adapt_hess_kernel at the headline size (source page needs -lineinfo: the AOT objects have it)
TAG=${1:-ahncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
AH_ONCE=1 timeout 800 ncu --set full --import-source on --clock-control none -k regex:mpx_adapt_hess_kernel -s 2 -c 1 -o $OUT/ahess -f python profiles/tools/ahess_trace.py > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
ncu -i $OUT/ahess.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i $OUT/ahess.ncu-rep --page source --csv > $OUT/source.csv 2>/dev/null
ls -la $OUT
