#!/bin/bash
mkdir -p gpurun_out/s11
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s11/pytest.log 2>&1
tail -4 gpurun_out/s11/pytest.log
timeout 600 python profiles/tools/adaptive_time.py > gpurun_out/s11/adaptive.txt 2> gpurun_out/s11/err
MPX_ADAPT_WCOL=0 timeout 600 python profiles/tools/adaptive_time.py > gpurun_out/s11/adaptive_gather.txt 2>> gpurun_out/s11/err
cat gpurun_out/s11/adaptive.txt; echo ---; cat gpurun_out/s11/adaptive_gather.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s11/bench20.json 2>> gpurun_out/s11/err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s11/bench20.json"))
print(round(d["ms_per_step"]*1e3,2), round(d["roofline"]["frac"],3), "stream", round(d["roofline"]["stream_launch_us"],2), "e2e", round(d["e2e"]["value"],1), {k:round(v["value"],1) for k,v in d["e2e_variants"].items()})
PY
tail -3 gpurun_out/s11/err
