"""Small fixed workload for ncu: the C2 kernel variant (TCP CHV, records + stats) on 2^20
trajectories, two runs (the first warms up).  Usage: python scripts/profile_run.py [model]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pdmp_b200 as P

model = sys.argv[1] if len(sys.argv) > 1 else "tcp"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
pol = int(os.environ.get("PDMP_TSTOP_POLICY", "0"))
prec = os.environ.get("PDMP_PREC", "fp64")
tol = dict(reltol=1e-5, abstol=1e-7) if prec == "fp32" else {}
if model == "tcp":
    pb = P.models.problem("tcp", tspan=(0.0, 200.0))
    xc0 = 0.5 + np.random.default_rng(2024).random((n, 1))
    ens = P.Ensemble(pb, P.CHVGPU("tcp", precision=prec), trajectories=n, n_jumps=1000, sample_times=[50.0, 100.0, 150.0, 200.0],
                     max_records=n * 160, xc0=xc0, tstop_policy=pol, xd_transport="auto", **tol)
elif model == "morris_lecar":
    pb = P.models.problem("morris_lecar")
    ens = P.Ensemble(pb, P.CHVGPU("morris_lecar", precision=prec), trajectories=n, n_jumps=2200, max_records=n * 800, **tol)
elif model == "tcp2d":
    pb = P.models.problem("tcp2d", tspan=(0.0, 10.0))
    ens = P.Ensemble(pb, P.CHVGPU("tcp2d", precision=prec), trajectories=n, n_jumps=2000, no_records=True,
                     sample_times=list(10.0 * np.arange(1, 17) / 16), hist=[(0, 0.0, 4.0), (1, 0.0, 4.0)], hist_bins=256, **tol)
elif model == "sir_rejection":
    pb = P.models.problem("sir_rejection")
    ens = P.Ensemble(pb, P.RejectionGPU("sir_rejection"), trajectories=n, n_jumps=1000, max_records=n * 260,
                     save_rate=bool(int(os.environ.get("PDMP_SAVE_RATE", "0"))))
if model == "neuron_mf":
    # RejectionExact (one warp per trajectory): one-shot solves, no resident handle
    pb = P.models.problem("neuron_mf")
    for seed in (0, 1):
        r = P.solve(pb, P.RejectionExactGPU("neuron_mf"), trajectories=n, n_jumps=1000, seed=seed, save_c_count=2, save_d_count=2)
        c = r.counters
        print(model, n, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in c.items()},
              "jumps/s", c["total_jumps"] / (c["kernel_ms"] * 1e-3), "candidates/s", c["total_candidates"] / (c["kernel_ms"] * 1e-3))
    sys.exit(0)
import time
for seed in (0, 1):
    t0 = time.perf_counter()
    ens.run(seed=seed)
    c = ens.counters()
    c["step_wall_ms"] = round(1e3 * (time.perf_counter() - t0), 3)
    print(model, n, "JT", os.environ.get("PDMP_JUMP_THRESH"), "RT", os.environ.get("PDMP_REFILL_THRESH"), "pol", pol, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in c.items()},
          "jumps/s", c["total_jumps"] / (c["kernel_ms"] * 1e-3))
