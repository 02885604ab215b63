#!/bin/bash
mkdir -p gpurun_out/s9
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s9/pytest.log 2>&1
tail -4 gpurun_out/s9/pytest.log
for i in 1 2; do
timeout 300 python bench.py --config 5 --steps 20 --warmup 5 --no-cpu --no-e2e-variants >> gpurun_out/s9/c5_fused.jsonl 2>> gpurun_out/s9/err
MPX_FUSE_PHASES=0 timeout 300 python bench.py --config 5 --steps 20 --warmup 5 --no-cpu --no-e2e-variants >> gpurun_out/s9/c5_unfused.jsonl 2>> gpurun_out/s9/err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e-variants >> gpurun_out/s9/head.jsonl 2>> gpurun_out/s9/err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e-variants --graph >> gpurun_out/s9/head_graph.jsonl 2>> gpurun_out/s9/err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e-variants >> gpurun_out/s9/head.jsonl 2>> gpurun_out/s9/err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e-variants --graph >> gpurun_out/s9/head_graph.jsonl 2>> gpurun_out/s9/err
python - <<'PY'
import json
for f in ("c5_fused","c5_unfused","head","head_graph"):
    for l in open(f"gpurun_out/s9/{f}.jsonl"):
        d=json.loads(l); print(f, round(d["ms_per_step"]*1e3,2), round(d["roofline"]["frac"],3), "iso", round(d["roofline"]["isolated_launch_us_median"],2), d["gpu_launches"], d["config"]["program"])
PY
tail -3 gpurun_out/s9/err
