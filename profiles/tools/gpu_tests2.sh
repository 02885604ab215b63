#!/bin/bash
mkdir -p gpurun_out/t2
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t2/pytest.log 2>&1
tail -8 gpurun_out/t2/pytest.log
