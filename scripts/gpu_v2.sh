#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sweep_v2.log
PDMP_KERNEL=1 PDMP_TSTOP_POLICY=1 timeout 300 python scripts/profile_run.py tcp 4194304 2>&1 | tail -1 >> gpurun_out/sweep_v2.log
for sw in 40 48 56 64; do for jb in 16 24 32; do
PDMP_SW=$sw PDMP_JB=$jb PDMP_TSTOP_POLICY=1 timeout 300 python scripts/profile_run.py tcp 4194304 2>&1 | tail -1 | sed "s/^/SW=$sw JB=$jb /" >> gpurun_out/sweep_v2.log
done; done
for m in morris_lecar tcp2d; do
PDMP_KERNEL=1 timeout 300 python scripts/profile_run.py $m 1048576 2>&1 | tail -1 >> gpurun_out/sweep_v2.log
timeout 300 python scripts/profile_run.py $m 1048576 2>&1 | tail -1 | sed "s/^/v2 /" >> gpurun_out/sweep_v2.log
done
sed -E "s/\{.*'kernel_ms'/kernel_ms/" gpurun_out/sweep_v2.log
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
