#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sweep5.log
run() { "$@" 2>&1 | tail -1 | sed -E "s/\{.*'kernel_ms'/kernel_ms/" >> gpurun_out/sweep5.log; }
PDMP_TSTOP_POLICY=1 run python scripts/profile_run.py tcp 4194304
run python scripts/profile_run.py tcp2d 2097152
run python scripts/profile_run.py morris_lecar 1048576
PDMP_JUMP_THRESH=4 run python scripts/profile_run.py morris_lecar 1048576
PDMP_JUMP_THRESH=8 run python scripts/profile_run.py morris_lecar 1048576
run python scripts/profile_run.py sir_rejection 1048576
cat gpurun_out/sweep5.log
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_1e7.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_1e7.log
python - <<'PY'
import json
for ln in open("gpurun_out/bench_1e7.log"):
    if ln.startswith("{"):
        d=json.loads(ln); print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"], d["roofline_output"]["achieved"], d["roofline_output"]["compact_ms_per_step"], d.get("cpu_baseline"))
    else: print(ln.strip()[:300])
PY
