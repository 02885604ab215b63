#!/bin/bash
mkdir -p gpurun_out/s8
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dynamic or sharded_objective" > gpurun_out/s8/pytest.log 2>&1
tail -3 gpurun_out/s8/pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/s8/bench20.json 2> gpurun_out/s8/err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s8/bench20.json"))
print(round(d["ms_per_step"]*1e3,2), round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), {k:round(v["value"],1) for k,v in d["e2e_variants"].items()})
PY
# steady-state DRAM traffic: caches NOT flushed between launches, 12 consecutive launches after 10 warm-ups
timeout 900 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum \
  -k regex:mpx_gjac2 -s 10 -c 12 --csv --log-file gpurun_out/s8/ncu_dram_steady.csv \
  python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e-variants > gpurun_out/s8/ncu_bench.log 2>&1
# every launch of a short bench run with its device time (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s8/ncu_launches.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e-variants > gpurun_out/s8/ncu_bench2.log 2>&1
# one full capture of the top kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpx_gjac2 -s 10 -c 2 -o gpurun_out/s8/prof_gjac2 \
  python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e-variants > gpurun_out/s8/ncu_bench3.log 2>&1
ls -la gpurun_out/s8
tail -3 gpurun_out/s8/err
