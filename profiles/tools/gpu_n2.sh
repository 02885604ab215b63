#!/bin/bash
# 2-GPU call: shim tests, then both bench arms under torchrun exactly as the driver launches them
TAG=${1:-n2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_shims.py tests/test_gpu_parity.py -q -m gpu -k "shims or peer or adaptive" ) > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; tail -3 $OUT/bench_n2.err; cut -c1-900 $OUT/bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $OUT/bench_ref_n2.json 2> $OUT/bench_ref_n2.err; tail -3 $OUT/bench_ref_n2.err; cut -c1-400 $OUT/bench_ref_n2.json
