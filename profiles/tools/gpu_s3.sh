#!/bin/bash
mkdir -p gpurun_out/s3
export MPX_TRACE_OUT=gpurun_out/s3
for i in 1 2 3; do timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu >> gpurun_out/s3/bench20.jsonl 2>> gpurun_out/s3/err; done
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 10 --no-cpu >> gpurun_out/s3/bench200.jsonl 2>> gpurun_out/s3/err
for c in 2 3 4 5; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu >> gpurun_out/s3/bench_configs.jsonl 2>> gpurun_out/s3/err; done
MPX_TRACE=1 timeout 300 python profiles/tools/trace_timeline.py --mode chain > gpurun_out/s3/trace_chain.json 2>> gpurun_out/s3/err
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s3/pytest.log 2>&1
tail -3 gpurun_out/s3/pytest.log
tail -3 gpurun_out/s3/err
