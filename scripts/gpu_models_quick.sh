#!/bin/bash
mkdir -p gpurun_out
( python scripts/profile_run.py morris_lecar 1048576 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-200
python scripts/profile_run.py tcp2d 4194304 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-200
python scripts/profile_run.py sir_rejection 2097152 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-200
PDMP_SAVE_RATE=1 python scripts/profile_run.py sir_rejection 2097152 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-200 ) | tee gpurun_out/models_quick.log
