#!/bin/bash
# round-1 evidence pass: parity tests, both bench arms, ncu launch list of the bench command,
# one --set full capture of the ensemble kernel and of the compaction kernel
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref.log
timeout 900 python bench.py > gpurun_out/bench_1e7.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_1e7.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_for_ncu_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_for_ncu.log 2>&1
echo "ncu-launches rc=$?"
python scripts/profile_run.py tcp > gpurun_out/profile_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -s 1 -c 1 -o gpurun_out/prof_tcp_r3 -f python scripts/profile_run.py tcp > gpurun_out/ncu_tcp_r3.log 2>&1
echo "ncu-full-ensemble rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:compact_kernel -s 1 -c 1 -o gpurun_out/prof_compact_r3 -f python scripts/profile_run.py tcp > gpurun_out/ncu_compact_r3.log 2>&1
echo "ncu-full-compact rc=$?"
for f in pytest_gpu bench_ref bench_1e7 profile_plain; do echo "== $f"; tail -n 4 gpurun_out/$f.log | cut -c1-600; done
