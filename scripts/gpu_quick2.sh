#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
PDMP_TSTOP_POLICY=1 timeout 300 python scripts/profile_run.py tcp 4194304 2>&1 | tail -1 | sed -E "s/\{.*'aux_attempts'/aux/"
for m in morris_lecar tcp2d sir_rejection; do timeout 300 python scripts/profile_run.py $m 1048576 2>&1 | tail -1 | sed -E "s/\{.*'aux_attempts'/aux/"; done
