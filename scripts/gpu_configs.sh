#!/bin/bash
# usage: gpu_configs.sh N   (N GPUs; N=1 runs without torchrun)
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1500 python scripts/run_configs.py > gpurun_out/configs_n1.jsonl 2> gpurun_out/configs_n1.err; echo "rc=$?" >> gpurun_out/configs_n1.err
else
  nvidia-smi -L > gpurun_out/multi_box_$N.txt; free -g | head -2 >> gpurun_out/multi_box_$N.txt; nproc >> gpurun_out/multi_box_$N.txt
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n$N.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/run_configs.py --no-parity > gpurun_out/configs_n$N.jsonl 2> gpurun_out/configs_n$N.err; echo "rc=$?" >> gpurun_out/configs_n$N.err
  grep -E "^\{|rc=" gpurun_out/bench_n$N.log | cut -c1-200
fi
cut -c1-400 gpurun_out/configs_n$N.jsonl; tail -5 gpurun_out/configs_n$N.err
