#!/bin/bash
TAG=${1:-ah1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1500 python -m pytest tests/test_gpu_adaptive.py tests/test_golden.py -x -q -m gpu ) > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log | cut -c1-200
timeout 600 python profiles/tools/adaptive_time.py 2>$OUT/err | tee $OUT/adaptive.txt | grep "hess" | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 12 --csv --log-file $OUT/launches.csv python profiles/tools/adaptive_time.py > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(l for l in open("$OUT/launches.csv") if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[1:]: print(r[ki][:60], r[vi])
PY
tail -3 $OUT/err
