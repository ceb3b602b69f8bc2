"""Generates tests/golden/closed_forms.npz — the portable known-answer vectors of the
reference's test-suite (closed-form samplers, oracle/closed_forms.py) on fixed uniform
streams.  The reference draws from Julia's global Xoshiro (unavailable here: no julia);
the closed forms are RNG-agnostic, so any fixed stream pins the same mathematics.

Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import closed_forms as cf  # noqa: E402


def stream(seed, n):
    u = np.random.default_rng(seed).random(n)
    return np.clip(u, 2.0 ** -53, 1.0 - 2.0 ** -53)


def main():
    out = {}
    # examples/tcp.jl: seed 43143, nj = 10 ; plus a long 700-jump run (jump times stall in double precision near jump 756 of this stream)
    for tag, nj in (("tcp10", 10), ("tcp700", 700)):
        u = stream(43143, 2 * 1000 + 8)
        t, x, d = cf.tcp_chv(u, nj=nj)
        out[f"{tag}_u"], out[f"{tag}_t"], out[f"{tag}_x"], out[f"{tag}_xd"] = u, t, x, d
    # test/pdmpStiff.jl: seed 8, ti = 0.322156, nj = 50
    u = stream(8, 128)
    t, x, d = cf.stiff_chv(u, nj=50)
    out["stiff_u"], out["stiff_t"], out["stiff_x"], out["stiff_xd"] = u, t, x, d
    # examples/pdmpExplosion.jl: seed 18, ti = 0.332, nj = 50 (pick the first stream >= 18 without NaN)
    seed = 18
    while True:
        u = stream(seed, 128)
        t, x, d = cf.explosion_chv(u, nj=50)
        if np.all(np.isfinite(t)) and np.all(x > 0):
            break
        seed += 1
    out["explosion_seed"] = np.array([seed])
    out["explosion_u"], out["explosion_t"], out["explosion_x"], out["explosion_xd"] = u, t, x, d
    # examples/tcp_rejection.jl: seed 1234, nj = 50
    u = stream(1234, 4096)
    t, x, d, nrej = cf.tcp_rejection(u, nj=50)
    out["tcprej_u"], out["tcprej_t"], out["tcprej_x"], out["tcprej_xd"] = u, t, x, d
    out["tcprej_nrej"] = np.array([nrej])
    # test/pdmpStiff.jl rejection: seed 8, nj = 4
    u = stream(8, 8192)
    t, x, d, nrej = cf.stiff_rejection(u, nj=4)
    out["stiffrej_u"], out["stiffrej_t"], out["stiffrej_x"], out["stiffrej_xd"] = u, t, x, d
    out["stiffrej_nrej"] = np.array([nrej])
    # examples/sir-rejection.jl: exact thinning, 6 trajectories, n_jumps = 1000, tf = 150
    U = np.stack([stream(1234 + k, 4096) for k in range(6)])
    out["sir_u"] = U
    for k in range(6):
        r = cf.sir_thinning(U[k])
        out[f"sir{k}_t"], out[f"sir{k}_xd"] = r["time"], r["xd"]
        out[f"sir{k}_meta"] = np.array([r["njumps"], r["nrejected"], r["status"], r["ndraws"]])
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "closed_forms.npz"), **out)
    print("wrote closed_forms.npz:", {k: v.shape for k, v in out.items() if not k.endswith("_u")})


if __name__ == "__main__":
    main()
