#!/bin/bash
mkdir -p gpurun_out/s2
export MPX_TRACE_OUT=gpurun_out/s2
MPX_TRACE=1 timeout 300 python profiles/tools/trace_timeline.py --mode chain > gpurun_out/s2/trace_chain.json 2>> gpurun_out/s2/err
MPX_TRACE=1 timeout 300 python profiles/tools/trace_timeline.py --mode isolated > gpurun_out/s2/trace_isolated.json 2>> gpurun_out/s2/err
MPX_TRACE=1 timeout 300 python profiles/tools/trace_timeline.py --mode chain --K 8192 --deg 20 --scheme LGL > gpurun_out/s2/trace_chain_c4.json 2>> gpurun_out/s2/err
MPX_TRACE=1 MPX_PDL=0 timeout 300 python profiles/tools/trace_timeline.py --mode chain > gpurun_out/s2/trace_chain_nopdl.json 2>> gpurun_out/s2/err
tail -3 gpurun_out/s2/err
