#!/bin/bash
# adaptive Hessian at the headline size: parity tests of that path, device time, per-segment timeline, ncu launch list
TAG=${1:-ah}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_adaptive.py tests/test_golden.py -x -q -m gpu -k "hess" ) > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log | cut -c1-200
timeout 600 python profiles/tools/ahess_trace.py 2>$OUT/err | tee $OUT/time.txt
MPX_TRACE=1 timeout 600 python profiles/tools/ahess_trace.py 2>>$OUT/err | tee $OUT/trace.txt
AH_ONCE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python profiles/tools/ahess_trace.py > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(l for l in open("$OUT/launches.csv") if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[-8:]: print(r[ki][:70], r[vi])
PY
tail -3 $OUT/err
