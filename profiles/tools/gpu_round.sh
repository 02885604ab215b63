#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list and one full capture of the top kernel.
# usage: gpurun --timeout 1200 -- bash profiles/tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $OUT/clocks.csv &
SMI=$!
nproc > $OUT/nproc.txt
( time timeout 600 python -m pytest tests -x -q -m gpu ) > $OUT/pytest_gpu.log 2>&1
tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
timeout 400 python bench.py > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
kill $SMI
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mpx_gjac -s 3 -c 2 -f -o $OUT/prof_gjac \
    python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1
ls -la $OUT
