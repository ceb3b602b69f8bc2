#!/bin/bash
# multi-GPU pass on N GPUs: the in-library sharding test, PCIe peaks at N ranks, bench through torchrun and in
# single-process mode (opts.devices).  usage: gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_multi_box_$N.txt
timeout 900 python -m pytest tests/test_gpu_v8.py -q -m gpu -k "multi_device or statistics_chunked" > gpurun_out/r2_multi_pytest_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_multi_pytest_$N.log
tail -n 3 gpurun_out/r2_multi_pytest_$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29510 scripts/pcie_probe.py > gpurun_out/r2_pcie_n$N.log 2>&1; tail -n 1 gpurun_out/r2_pcie_n$N.log | cut -c1-300
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-policy0 > gpurun_out/r2_bench_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2_bench_n$N.log
grep -E "^\{" gpurun_out/r2_bench_n$N.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('torchrun', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('roofline_e2e',{}).get('frac'))"
timeout 1200 python bench.py --gpus $N --steps 5 --warmup 3 --single-process > gpurun_out/r2_bench_sp_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2_bench_sp_n$N.log
grep -E "^\{" gpurun_out/r2_bench_sp_n$N.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('single-process', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
tail -n 2 gpurun_out/r2_bench_sp_n$N.log | cut -c1-300
if [ "$N" -ge 4 ]; then
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/run_configs.py --no-parity > gpurun_out/r2_configs_n$N.jsonl 2> gpurun_out/r2_configs_n$N.err; echo "configs rc=$?"
  cut -c1-330 gpurun_out/r2_configs_n$N.jsonl
fi
