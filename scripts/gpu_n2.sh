#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 900 -x -k "ordered or resident or sharding" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/run_configs.py --no-parity --only C5 --scale 0.02 2>&1 | grep -E "^\{|rror" | cut -c1-900
timeout 900 python scripts/run_configs.py --no-parity --only C5 --scale 0.02 2>&1 | grep -E "^\{|rror" | cut -c1-900
