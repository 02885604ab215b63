#!/bin/bash
# round 2, session 1: sanity + baseline numbers with the gate kernel + K2 timelines
mkdir -p gpurun_out/s1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/s1/clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s1/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1/pytest.log
for i in 1 2 3; do timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu >> gpurun_out/s1/bench20.jsonl 2>> gpurun_out/s1/bench.err; done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-gate >> gpurun_out/s1/bench20_nogate.jsonl 2>> gpurun_out/s1/bench.err
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 10 --no-cpu >> gpurun_out/s1/bench200.jsonl 2>> gpurun_out/s1/bench.err
for c in 2 3 4 5; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu >> gpurun_out/s1/bench_configs.jsonl 2>> gpurun_out/s1/bench.err; done
MPX_TRACE=1 timeout 300 python profiles/tools/trace_timeline.py --mode chain > gpurun_out/s1/trace_chain.json 2>> gpurun_out/s1/bench.err
MPX_TRACE=1 timeout 300 python profiles/tools/trace_timeline.py --mode isolated > gpurun_out/s1/trace_isolated.json 2>> gpurun_out/s1/bench.err
MPX_TRACE=1 timeout 300 python profiles/tools/trace_timeline.py --mode chain --K 8192 --deg 20 --scheme LGL > gpurun_out/s1/trace_chain_c4.json 2>> gpurun_out/s1/bench.err
kill $SMI
tail -3 gpurun_out/s1/pytest.log; cat gpurun_out/s1/bench20.jsonl | cut -c1-400; tail -5 gpurun_out/s1/bench.err
