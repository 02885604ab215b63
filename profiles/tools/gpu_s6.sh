#!/bin/bash
mkdir -p gpurun_out/s6
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dynamic" > gpurun_out/s6/pytest_dyn.log 2>&1
tail -5 gpurun_out/s6/pytest_dyn.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/s6/bench20.json 2> gpurun_out/s6/err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/s6/bench_ref.json 2>> gpurun_out/s6/err
for c in 2 3 4 5; do timeout 600 python bench.py --config $c --steps 20 --warmup 5 >> gpurun_out/s6/bench_configs.jsonl 2>> gpurun_out/s6/err; done
nproc > gpurun_out/s6/host.txt; lscpu | head -30 >> gpurun_out/s6/host.txt; numactl -H >> gpurun_out/s6/host.txt 2>&1; nvidia-smi topo -m >> gpurun_out/s6/host.txt 2>&1
tail -5 gpurun_out/s6/err
cat gpurun_out/s6/bench20.json | cut -c1-300
