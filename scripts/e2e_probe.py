"""End-to-end step (host buffers in, every saved point out) of the C2 workload for a few chunk sizes:
where the wall time goes next to the pinned-copy lower bound.  Usage: python scripts/e2e_probe.py [n_traj]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
chunks = [int(c) for c in os.environ.get("CHUNKS", "1048576").split(",")]
for ch in chunks:
    os.environ["PDMP_CHUNK"] = str(ch)
    import pdmp_b200 as P
    pb = P.models.problem("tcp", tspan=(0.0, 200.0))
    ics = 0.5 + np.random.default_rng(2024).random((n, 1))
    ens = P.Ensemble(pb, P.CHVGPU("tcp", ode="dp5"), trajectories=n, n_jumps=1000, reltol=1e-7, abstol=1e-9,
                     sample_times=[50.0, 100.0, 150.0, 200.0], xd_transport="auto", tstop_policy=1, xc0=ics, max_records=int(n * 160),
                     save_positions=(False, True))
    r = ens.run_host(seed=999, xc0=ics); del r
    for s in range(3):
        t0 = time.perf_counter()
        r = ens.run_host(seed=s, xc0=ics)
        dt = time.perf_counter() - t0
        c = r.counters
        nbytes = r.time.nbytes + r.xc.nbytes + r._xd_raw.nbytes + r.offsets.nbytes + r.njumps.nbytes + r.nrejected.nbytes + r.status.nbytes
        print(f"chunk {ch} step {s}: wall {1e3 * dt:.1f} ms, {nbytes / 1e9:.2f} GB -> {nbytes / dt / 1e9:.1f} GB/s; kernel {c['kernel_ms']:.1f} compact {c['compact_ms']:.1f} "
              f"d2h {c['d2h_ms']:.1f} h2d {c['h2d_ms']:.2f} ms; jumps/s {c['total_jumps'] / dt:.4g}", flush=True)
        del r
    ens.close()
