#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1e7.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_1e7.log
python scripts/profile_run.py tcp > gpurun_out/profile_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -s 1 -c 1 -o gpurun_out/prof_tcp_r1 -f python scripts/profile_run.py tcp > gpurun_out/ncu_tcp.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu_tcp.log
for f in smoke pytest_gpu bench_1e7 profile_plain ncu_tcp; do echo "== $f"; tail -n 6 gpurun_out/$f.log; done
