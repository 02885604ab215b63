#!/bin/bash
mkdir -p gpurun_out/t
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t/pytest.log 2>&1
tail -15 gpurun_out/t/pytest.log
