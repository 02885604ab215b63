#!/bin/bash
mkdir -p gpurun_out/s7
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s7/pytest.log 2>&1
tail -5 gpurun_out/s7/pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/s7/bench20.json 2> gpurun_out/s7/err
for t in 4 8; do MPX_HOST_THREADS=$t timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/s7/bench20_t$t.json 2>> gpurun_out/s7/err; done
tail -5 gpurun_out/s7/err
python - <<'PY'
import json
for f in ("bench20","bench20_t4","bench20_t8"):
    d=json.load(open(f"gpurun_out/s7/{f}.json"))
    print(f, round(d["ms_per_step"]*1e3,2), round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), {k:round(v["value"],1) for k,v in d["e2e_variants"].items()})
PY
