#!/bin/bash
OUT=gpurun_out/${1:-sp1}
mkdir -p $OUT
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu ) > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log | cut -c1-200
for s in 1 0 1 0; do
for c in 1 3 4 5; do
  if [ $c = 1 ]; then CF=""; else CF="--config $c"; fi
  MPX_V2_SPREAD=$s timeout 300 python bench.py $CF --steps 20 --warmup 5 --no-cpu --no-e2e-variants 2>>$OUT/err > $OUT/t.json
  python - <<PY
import json
d=json.load(open("$OUT/t.json")); r=d["roofline"]
print("spread=$s config $c us", round(d["ms_per_step"]*1e3,2), "frac", round(r["frac"],3), "stream", round(r["stream_launch_us"],2))
PY
done; done
tail -3 $OUT/err
