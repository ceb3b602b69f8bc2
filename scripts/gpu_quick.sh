#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
PDMP_TSTOP_POLICY=1 timeout 300 python scripts/profile_run.py tcp 4194304 2>&1 | tail -1 | sed -E "s/\{.*'kernel_ms'/kernel_ms/"
for m in morris_lecar tcp2d sir_rejection; do timeout 300 python scripts/profile_run.py $m 1048576 2>&1 | tail -1 | sed -E "s/\{.*'kernel_ms'/kernel_ms/"; done
timeout 900 python bench.py --no-cpu > gpurun_out/bench_1e7.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_1e7.log
python - <<'PY'
import json
for ln in open("gpurun_out/bench_1e7.log"):
    if ln.startswith("{"):
        d=json.loads(ln); print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"], d["roofline_output"]["achieved"], d["roofline_output"]["compact_ms_per_step"])
    else: print(ln.strip()[:300])
PY
