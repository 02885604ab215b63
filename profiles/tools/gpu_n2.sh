#!/bin/bash
# 2-GPU call: both bench arms under torchrun exactly as the driver launches them; stdout must be ONE JSON line
TAG=${1:-n2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "stdout lines: $(wc -l < $OUT/bench_n2.json)"; cut -c1-200 $OUT/bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $OUT/bench_ref_n2.json 2> $OUT/bench_ref_n2.err; echo "stdout lines: $(wc -l < $OUT/bench_ref_n2.json)"; cut -c1-200 $OUT/bench_ref_n2.json
timeout 300 python bench.py --no-cpu --steps 50 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "stdout lines: $(wc -l < $OUT/bench_n1.json)"; cut -c1-200 $OUT/bench_n1.json
