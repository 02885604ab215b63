#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out/n$N
nvidia-smi topo -m > gpurun_out/n$N/topo.txt 2>&1; nproc >> gpurun_out/n$N/topo.txt; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" >> gpurun_out/n$N/topo.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n$N/bench.json 2> gpurun_out/n$N/err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --config 4 --gpus $N --steps 20 --warmup 5 > gpurun_out/n$N/bench_c4.json 2>> gpurun_out/n$N/err
tail -5 gpurun_out/n$N/err
python - <<PY
import json
for f in ("bench","bench_c4"):
    try:
        d=json.load(open("gpurun_out/n$N/%s.json"%f))
    except Exception as e:
        print(f, "no json", e); continue
    print(f, "value", round(d["value"],1), d["scaling"], "ms", round(d["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "e2e ms", round(d["e2e"]["ms_per_step"],2), "copy-only ms", d["e2e"].get("d2h_copy_only_ms_per_step"))
    for k in ("allgather","allgather_nccl"):
        a=d.get(k)
        if a: print("  ",k, {x:(round(v,3) if isinstance(v,float) else v) for x,v in a.items() if x not in ("how","note")})
PY
