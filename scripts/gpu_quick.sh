#!/bin/bash
# quick A/B pass: GPU parity tests, then the C2 kernel at 4 M and 10 M trajectories (both step policies)
# usage: bash scripts/gpu_quick.sh TAG [notest]
TAG=${1:-q}
mkdir -p gpurun_out
if [ "$2" != "notest" ]; then
  timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
  tail -n 6 gpurun_out/${TAG}_pytest.log | cut -c1-400
fi
for pol in 0 1; do
  PDMP_TSTOP_POLICY=$pol timeout 300 python scripts/profile_run.py tcp 4194304 2>&1 | tail -n 1 | cut -c1-420
done | tee gpurun_out/${TAG}_4M.log
PDMP_TSTOP_POLICY=1 timeout 300 python scripts/profile_run.py tcp 10000000 2>&1 | tail -n 1 | cut -c1-420 | tee gpurun_out/${TAG}_10M.log
