#!/bin/bash
# GPU parity suite only; usage: gpu_tests.sh TAG [pytest args]
TAG=${1:-t}; shift
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 "$@" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -n 60 gpurun_out/${TAG}_pytest.log | cut -c1-300
