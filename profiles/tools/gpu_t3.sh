#!/bin/bash
mkdir -p gpurun_out/t3
timeout 900 python -m pytest tests/test_gpu_adaptive.py -m gpu -q -x > gpurun_out/t3/pytest_adapt.log 2>&1
tail -25 gpurun_out/t3/pytest_adapt.log
