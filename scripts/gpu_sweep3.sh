#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sweep3.log
for jt in 8 12 16 20; do PDMP_JUMP_THRESH=$jt PDMP_TSTOP_POLICY=1 python scripts/profile_run.py tcp 4194304 2>&1 | tail -1 | sed -E "s/\{.*'kernel_ms'/kernel_ms/" >> gpurun_out/sweep3.log; done
for jt in 8 12 16 20; do PDMP_JUMP_THRESH=$jt python scripts/profile_run.py tcp2d 2097152 2>&1 | tail -1 | sed -E "s/\{.*'kernel_ms'/kernel_ms/" >> gpurun_out/sweep3.log; done
for rt in 1 4 8 12; do PDMP_REFILL_THRESH=$rt python scripts/profile_run.py tcp2d 2097152 2>&1 | tail -1 | sed -E "s/\{.*'kernel_ms'/kernel_ms/" >> gpurun_out/sweep3.log; done
for jt in 8 16 20; do PDMP_JUMP_THRESH=$jt python scripts/profile_run.py morris_lecar 1048576 2>&1 | tail -1 | sed -E "s/\{.*'kernel_ms'/kernel_ms/" >> gpurun_out/sweep3.log; done
cat gpurun_out/sweep3.log
