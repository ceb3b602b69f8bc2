run() { "$@" 2>&1 | tail -1 | sed -E "s/\{.*'kernel_ms'/kernel_ms/"; }
PDMP_TSTOP_POLICY=1 run python scripts/profile_run.py tcp 4194304
run python scripts/profile_run.py tcp2d 2097152
run python scripts/profile_run.py morris_lecar 1048576
run python scripts/profile_run.py sir_rejection 2097152
python scripts/profile_run.py neuron_mf 65536 2>&1 | tail -1 | cut -c1-20,120-500
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -3
