#!/bin/bash
# One gpurun call: GPU parity tests, smoke, the bench line (+ reference arm), ablations, the ncu launch list and one
# full capture of the top kernel.
# usage: gpurun --timeout 1500 -- bash profiles/tools/gpu_round.sh [tag]
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
for cfg in "MPX_PDL=0" "MPX_CONST_FIRST=0" "MPX_V2_NBUF=1" "MPX_NOSPEC=1" "MPX_JIT=-1" "MPX_KERNEL=v4" "MPX_KERNEL=v1"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --no-cpu --steps 200 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(json.dumps({'program': d['config']['program'], 'value': round(d['value']), 'us_per_step': round(d['ms_per_step']*1e3,2), 'frac': round(d['roofline']['frac'],3), 'isolated_us': round(d['roofline']['isolated_launch_us_median'],2)}))
"
done > $OUT/ablations.txt 2>&1
cat $OUT/ablations.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mpx_gjac -s 3 -c 2 -f -o $OUT/prof_gjac \
    python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1
ls -la $OUT
