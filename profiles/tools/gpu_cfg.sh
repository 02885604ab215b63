#!/bin/bash
# parity of the g + jac kernels, then one bench line per configuration: usage gpu_cfg.sh <tag> [env assignments...]
TAG=${1:-cfg}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_adaptive.py -x -q -m gpu ) > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log | cut -c1-200
: > $OUT/configs.jsonl
for c in 1 2 3 4 5; do
  if [ $c = 1 ]; then CF=""; else CF="--config $c"; fi
  env "$@" timeout 300 python bench.py $CF --steps 20 --warmup 5 --no-cpu 2>>$OUT/err >> $OUT/configs.jsonl
done
python - <<PY
import json
for l in open("$OUT/configs.jsonl"):
    d=json.loads(l); r=d["roofline"]
    print(d["metric"][-45:], "| us", round(d["ms_per_step"]*1e3,2), "frac", round(r["frac"],3), "stream", round(r["stream_launch_us"],2), "e2e", round(d["e2e"]["value"],1))
PY
tail -3 $OUT/err
