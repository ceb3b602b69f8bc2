#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sweep_v2.log
PDMP_KERNEL=1 PDMP_TSTOP_POLICY=1 timeout 300 python scripts/profile_run.py tcp 4194304 2>&1 | tail -1 >> gpurun_out/sweep_v2.log
for sw in 48 56 64; do for jb in 20 26 32; do
PDMP_SW=$sw PDMP_JB=$jb PDMP_TSTOP_POLICY=1 timeout 300 python scripts/profile_run.py tcp 4194304 2>&1 | tail -1 | sed "s/^/SW=$sw JB=$jb /" >> gpurun_out/sweep_v2.log
done; done
sed -E "s/\{.*'kernel_ms'/kernel_ms/" gpurun_out/sweep_v2.log
