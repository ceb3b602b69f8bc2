#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sweep4.log
run() { "$@" 2>&1 | tail -1 | sed -E "s/\{.*'kernel_ms'/kernel_ms/" >> gpurun_out/sweep4.log; }
PDMP_TSTOP_POLICY=1 run python scripts/profile_run.py tcp 4194304
PDMP_JUMP_THRESH=16 PDMP_REFILL_THRESH=2 PDMP_TSTOP_POLICY=1 run python scripts/profile_run.py tcp 4194304
run python scripts/profile_run.py tcp2d 2097152
PDMP_JUMP_THRESH=16 PDMP_REFILL_THRESH=8 run python scripts/profile_run.py tcp2d 2097152
run python scripts/profile_run.py morris_lecar 1048576
PDMP_JUMP_THRESH=8 PDMP_REFILL_THRESH=2 run python scripts/profile_run.py morris_lecar 1048576
PDMP_JUMP_THRESH=4 PDMP_REFILL_THRESH=2 run python scripts/profile_run.py morris_lecar 1048576
cat gpurun_out/sweep4.log
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -3
