#!/bin/bash
# usage: gpu_multi.sh N
N=$1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_box_$N.txt; free -g | head -2 >> gpurun_out/multi_box_$N.txt; nproc >> gpurun_out/multi_box_$N.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n$N.log
cat gpurun_out/multi_box_$N.txt | tail -3
python - <<PY
import json
for ln in open("gpurun_out/bench_n$N.log"):
    if ln.startswith("{"):
        d=json.loads(ln); print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"], d["roofline"]["frac"], d["clocks"])
    elif "rror" in ln or "rc=" in ln: print(ln.strip()[:300])
PY
