#!/bin/bash
# usage: gpu_k.sh TAG "pytest -k expression"
OUT=gpurun_out/$1
mkdir -p $OUT
( time timeout 900 python -m pytest tests -q -m gpu -k "$2" ) > $OUT/pytest.log 2>&1; tail -25 $OUT/pytest.log | cut -c1-220
