#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
PDMP_NO_SHARED_POOL=1 PDMP_CHUNK=8192 timeout 600 python -m pytest tests -q -m gpu --timeout 900 -x -k "chunked or resident or large or capacity" 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_1e7.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_1e7.log
python - <<'PY'
import json
for ln in open("gpurun_out/bench_1e7.log"):
    if ln.startswith("{"):
        d=json.loads(ln); print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"], d["roofline_output"])
    else: print(ln.strip()[:300])
PY
