#!/bin/bash
# first contact with the B200: probe the box, smoke, parity tests, a small bench
mkdir -p gpurun_out
{
echo "== box =="; nproc; free -g | head -2; cat /sys/fs/cgroup/memory.max 2>/dev/null; nvidia-smi -L; nvidia-smi --query-gpu=memory.total,memory.used --format=csv
which julia || echo "no julia"
} > gpurun_out/box.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --traj 1000000 --steps 3 --warmup 3 --cpu-sample 50000 > gpurun_out/bench_1e6.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_1e6.log
tail -5 gpurun_out/box.txt gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench_1e6.log
