#!/bin/bash
# resident step at 10^7 trajectories: one kernel vs 2 / 3 / 4 chunks whose compaction overlaps the next chunk's kernel
mkdir -p gpurun_out
( echo "single"; PDMP_TSTOP_POLICY=1 python scripts/profile_run.py tcp 10000000 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-200
for ch in 5000000 3400000 2500000; do echo "chunk $ch"; PDMP_NO_SHARED_POOL=1 PDMP_CHUNK=$ch PDMP_TSTOP_POLICY=1 python scripts/profile_run.py tcp 10000000 2>&1 | tail -n 1 | sed -e 's/{.*kernel_ms/kernel_ms/' | cut -c1-200; done ) | tee gpurun_out/chunks.log
