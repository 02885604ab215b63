#!/bin/bash
# ncu evidence of the shipped g + jac_g kernel: steady-state DRAM counters, launch list, one full capture
OUT=gpurun_out/${1:-ncur}
mkdir -p $OUT
timeout 900 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum \
  -k regex:mpx_gjac2 -s 10 -c 12 --csv --log-file $OUT/ncu_dram_steady.csv \
  python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e-variants > $OUT/ncu_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/ncu_launches.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e-variants > $OUT/ncu_bench2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpx_gjac2 -s 10 -c 2 -f -o $OUT/prof_gjac2 \
  python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e-variants > $OUT/ncu_bench3.log 2>&1
ls -la $OUT | head -12
