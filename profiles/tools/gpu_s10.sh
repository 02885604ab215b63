#!/bin/bash
mkdir -p gpurun_out/s10
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s10/pytest.log 2>&1
tail -4 gpurun_out/s10/pytest.log
timeout 300 python profiles/tools/evaluators_time.py > gpurun_out/s10/evaluators.txt 2> gpurun_out/s10/err
MPX_HESS_NOSTAGE=1 timeout 300 python profiles/tools/evaluators_time.py > gpurun_out/s10/evaluators_hess_nostage.txt 2>> gpurun_out/s10/err
cat gpurun_out/s10/evaluators.txt; grep hess gpurun_out/s10/evaluators_hess_nostage.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/s10/bench20.json 2>> gpurun_out/s10/err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/s10/bench_ref.json 2>> gpurun_out/s10/err
cut -c1-700 gpurun_out/s10/bench20.json
tail -3 gpurun_out/s10/err
