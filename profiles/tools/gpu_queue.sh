#!/bin/bash
TAG=${1:-q1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu ) > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log | cut -c1-200
for q in 1 0 1 0; do
  MPX_QUEUE=$q timeout 300 python bench.py --config 4 --steps 20 --warmup 5 --no-cpu 2>>$OUT/err > $OUT/c4_q$q.json
  python - <<PY
import json
d=json.load(open("$OUT/c4_q$q.json")); r=d["roofline"]
print("config4 queue=$q us", round(d["ms_per_step"]*1e3,2), "frac", round(r["frac"],3), "stream", round(r["stream_launch_us"],2))
PY
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu 2>>$OUT/err > $OUT/head.json
python - <<PY
import json
d=json.load(open("$OUT/head.json")); r=d["roofline"]
print("headline us", round(d["ms_per_step"]*1e3,2), "frac", round(r["frac"],3), "stream", round(r["stream_launch_us"],2))
PY
tail -3 $OUT/err
