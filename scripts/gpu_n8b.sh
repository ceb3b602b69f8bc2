#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_8.txt 2>&1
numactl -H >> gpurun_out/topo_8.txt 2>&1 || lscpu | grep -i numa >> gpurun_out/topo_8.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n8.log
python - <<'PY'
import json
for ln in open("gpurun_out/bench_n8.log"):
    if ln.startswith("{"):
        d=json.loads(ln); print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"], d.get("host_affinity_rank0"))
    elif "rror" in ln or "rc=" in ln: print(ln.strip()[:300])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scripts/run_configs.py --no-parity --only C5 2>&1 | grep "^{" | cut -c1-1200
