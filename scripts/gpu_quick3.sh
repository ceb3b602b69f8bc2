#!/bin/bash
mkdir -p gpurun_out
run() { "$@" 2>&1 | tail -1 | sed -E "s/\{.*'kernel_ms'/kernel_ms/"; }
PDMP_TSTOP_POLICY=1 run python scripts/profile_run.py tcp 4194304
run python scripts/profile_run.py tcp2d 2097152
run python scripts/profile_run.py morris_lecar 1048576
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -3
