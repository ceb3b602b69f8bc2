#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
: > gpurun_out/sweep.log
python scripts/profile_run.py tcp 4194304 2>&1 | tail -1 >> gpurun_out/sweep.log
PDMP_TSTOP_POLICY=1 python scripts/profile_run.py tcp 4194304 2>&1 | tail -1 >> gpurun_out/sweep.log
for m in morris_lecar tcp2d sir_rejection; do python scripts/profile_run.py $m 1048576 2>&1 | tail -1 >> gpurun_out/sweep.log; done
tail -n 5 gpurun_out/pytest_gpu.log; cat gpurun_out/sweep.log
