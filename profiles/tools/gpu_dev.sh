#!/bin/bash
# Development call: GPU parity tests, then bench lines for kernel variants (no CPU leg), then the write-bandwidth probe.
TAG=${1:-dev}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -x -q -m gpu ) > $OUT/pytest_gpu.log 2>&1
tail -15 $OUT/pytest_gpu.log
for cfg in "v2" "v2 MPX_JIT=-1" "v2 MPX_NOSPEC=1"; do
  set -- $cfg
  echo "== $cfg"
  env MPX_KERNEL=$1 $2 $3 timeout 300 python bench.py --no-cpu --steps 200 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(d['config']['program'], 'value', round(d['value']), 'us/step', round(d['ms_per_step']*1e3,2), 'isolated_us', round(d['roofline']['isolated_launch_us_median'],2), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
"
done 2>&1 | tee $OUT/variants.log

