#!/bin/bash
# round 2: the evidence run on one GPU -- full GPU test suite, smoke, the driver's bench command lines, every configuration
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r2/clocks.csv &
SMI=$!
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2/pytest_gpu.log 2>&1
tail -3 gpurun_out/r2/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke.log 2>&1; tail -4 gpurun_out/r2/smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench_reference_arm.json 2> gpurun_out/r2/err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench.json 2>> gpurun_out/r2/err
for c in 2 3 4 5; do timeout 600 python bench.py --config $c --steps 20 --warmup 5 >> gpurun_out/r2/bench_configs.jsonl 2>> gpurun_out/r2/err; done
timeout 300 python profiles/tools/evaluators_time.py > gpurun_out/r2/evaluators.txt 2>> gpurun_out/r2/err
kill $SMI
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2/bench.json")); r=json.load(open("gpurun_out/r2/bench_reference_arm.json"))
print("headline", round(d["value"]), "frac", round(d["roofline"]["frac"],3), "stream_us", round(d["roofline"]["stream_launch_us"],2), "e2e", round(d["e2e"]["value"],1), {k:round(v["value"],1) for k,v in d["e2e_variants"].items()}, "cpu", round(d["cpu_baseline"]["value"],1), "ref arm", round(r["value"],1))
for l in open("gpurun_out/r2/bench_configs.jsonl"):
    c=json.loads(l); print(c["metric"][-34:], round(c["ms_per_step"]*1e3,2), round(c["roofline"]["frac"],3), "stream", round(c["roofline"]["stream_launch_us"],2), "e2e", round(c["e2e"]["value"],1), "cpu", round(c["cpu_baseline"]["value"],1), c["cpu_baseline"]["cores"], c["gpu_launches"])
PY
tail -3 gpurun_out/r2/err
