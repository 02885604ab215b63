#!/bin/bash
mkdir -p gpurun_out/s5
export MPX_TRACE_OUT=gpurun_out/s5
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s5/pytest.log 2>&1
tail -3 gpurun_out/s5/pytest.log
for i in 1 2 3; do
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu >> gpurun_out/s5/bench20.jsonl 2>> gpurun_out/s5/err
MPX_CONST_QUEUE=0 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu >> gpurun_out/s5/bench20_noqueue.jsonl 2>> gpurun_out/s5/err
done
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 10 --no-cpu >> gpurun_out/s5/bench200.jsonl 2>> gpurun_out/s5/err
timeout 300 python bench.py --config 4 --steps 20 --warmup 5 --no-cpu >> gpurun_out/s5/bench_c4.jsonl 2>> gpurun_out/s5/err
MPX_CONST_QUEUE=0 timeout 300 python bench.py --config 4 --steps 20 --warmup 5 --no-cpu >> gpurun_out/s5/bench_c4.jsonl 2>> gpurun_out/s5/err
MPX_TRACE=1 timeout 300 python profiles/tools/trace_timeline.py --mode chain > gpurun_out/s5/trace_chain.json 2>> gpurun_out/s5/err
tail -3 gpurun_out/s5/err
