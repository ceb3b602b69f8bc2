"""Closed-form (analytical) samplers of the reference's known-answer tests.

TEST INFRASTRUCTURE (see oracle/pdmp_oracle.cpp header).  These are the portable golden
vectors of the reference: its test-suite pins the CHV / Rejection drivers by comparing
`solve(...)` against analytical samplers that consume the SAME uniform stream, so they are
RNG- and integrator-agnostic.  Each function restates one sampler and consumes `u`
(a 1-D array of uniforms in (0,1)) in exactly the order the Julia code calls `rand()`.

  tcp_chv            examples/tcp.jl:5-23             (test/runtests.jl:9-12,   tol 1e-4)
  stiff_chv          test/pdmpStiff.jl:5-27           (test/pdmpStiff.jl:119-127, 3e-3 / 4e-6)
  explosion_chv      examples/pdmpExplosion.jl:5-26   (test/runtests.jl:14-17,  tol 1e-4)
  tcp_rejection      examples/tcp_rejection.jl:27-65  (test/runtests.jl:41-44,  tol 1e-5)
  stiff_rejection    test/pdmpStiff.jl:29-69          (test/pdmpStiff.jl:172-179, tol 0.0043)

plus `sir_thinning`, an exact (ODE-free) restatement of rejection_diffeq! for the
zero-flow SIR model (examples/sir-rejection.jl), used to pin the discrete bookkeeping
(draw order, pfsample, nu update, counters) bit-exactly.
"""
import math

import numpy as np


class _Stream:
    def __init__(self, u):
        self.u = np.asarray(u, dtype=np.float64)
        self.k = 0

    def rand(self):
        v = float(self.u[self.k])
        self.k += 1
        return v


def tcp_chv(u, xc0=1.0, xd0=0, ti=0.0, nj=10):
    """examples/tcp.jl:5-23.  Returns (times, x, xd1) with x the TRUE continuous state
    (the Julia sampler stores 1/x in `xch`; only `th` is asserted there)."""
    r = _Stream(u)
    th, xh, dh = [ti], [xc0], [xd0]
    t, x, xd = ti, xc0, xd0
    while len(th) < nj:
        S = -math.log(r.rand())
        if xd % 2 == 0:      # F = +1, rate 1/x : x grows
            t += x * (math.exp(S) - 1.0)
            x = x * math.exp(S)
        else:                # F = -1
            t += x * (1.0 - math.exp(-S))
            x = x * math.exp(-S)
        xd += 1
        th.append(t); xh.append(x); dh.append(xd)
        r.rand()             # burns the pfsample draw (examples/tcp.jl:20)
    return np.array(th), np.array(xh), np.array(dh)


def stiff_chv(u, xc0=1.0, xd0=0, ti=0.322156, nj=50):
    """test/pdmpStiff.jl:5-27 (AnalyticalSampleCHV)."""
    r = _Stream(u)
    th, xh, dh = [ti], [xc0], [xd0]
    t, x, xd = ti, xc0, xd0
    while len(th) < nj:
        S = -math.log(r.rand())
        if xd % 2 == 0:
            t += 1.0 / 10.0 * math.log(1.0 + 10.0 * S / x)
            x = x + 10.0 * S
        else:
            t += 1.0 / (3.0 * x) * (math.exp(3.0 * S) - 1.0)
            x = x * math.exp(-3.0 * S)
        xd += 1
        th.append(t); xh.append(x); dh.append(xd)
        r.rand()
    return np.array(th), np.array(xh), np.array(dh)


def explosion_chv(u, xc0=1.0, xd0=0, ti=0.332, nj=50, rr=10.0):
    """examples/pdmpExplosion.jl:5-26.  NaN once 10 S > x in the contracting phase."""
    r = _Stream(u)
    th, xh, dh = [ti], [xc0], [xd0]
    t, x, xd = ti, xc0, xd0
    while len(th) < nj:
        S = -math.log(r.rand())
        a = -rr * (2 * (xd % 2) - 1)
        arg = a * S / x + 1.0
        dt = math.log(arg) / a if arg > 0 else float("nan")
        t += dt
        x = x + a * S
        xd += 1
        th.append(t); xh.append(x); dh.append(xd)
        r.rand()
    return np.array(th), np.array(xh), np.array(dh)


def _thinning(u, xc0, xd0, ti, nj, flow_even, flow_odd, rate_fn, bound_fn):
    r = _Stream(u)
    th, xh, dh = [ti], [xc0], [xd0]
    t, x = ti, xc0
    njumps, nrej = 1, 0
    lam = bound_fn(x)
    S = -math.log(r.rand()) / lam
    while njumps < nj:
        xd = dh[-1]
        t += S
        x = flow_even(x, S) if xd % 2 == 0 else flow_odd(x, S)
        lam = bound_fn(x)
        rate = rate_fn(x)
        reject = r.rand() < (1.0 - rate / lam)
        S = -math.log(r.rand()) / lam
        if not reject:
            th.append(t); xh.append(x); dh.append(xd + 1)
            njumps += 1
            r.rand()         # emulates pfsample
        else:
            nrej += 1
    return np.array(th), np.array(xh), np.array(dh), nrej


def tcp_rejection(u, xc0=0.0, xd0=0, ti=0.0, nj=50):
    """examples/tcp_rejection.jl:27-65: F = 1 | -10x, R = 1/(1+exp(-x)), bound 1."""
    return _thinning(u, xc0, xd0, ti, nj,
                     lambda x, S: x + S,
                     lambda x, S: x * math.exp(-10.0 * S),
                     lambda x: 1.0 / (1.0 + math.exp(-x)),
                     lambda x: 1.0)


def stiff_rejection(u, xc0=1.0, xd0=0, ti=0.322156, nj=4, p1=0.0):
    """test/pdmpStiff.jl:29-69 (AnalyticalSampleRejection): F = 10x | -3x^2, R = x + p1,
    bound p1 + 100."""
    return _thinning(u, xc0, xd0, ti, nj,
                     lambda x, S: x * math.exp(10.0 * S),
                     lambda x, S: x / (3.0 * S * x + 1.0),
                     lambda x: x + p1,
                     lambda x: p1 + 100.0)


def pfsample(w, U):
    """reference src/utils.jl:46-57 (1-based index)."""
    s = w[0]
    for v in w[1:]:
        s = s + v
    thr = U * s
    i, cw = 1, w[0]
    while cw < thr and i < len(w):
        i += 1
        cw += w[i - 1]
    return i


def sir_thinning(u, xd0=(99, 10, 0, 0), parms=(0.001, 0.01), ti=0.0, tf=150.0, n_jumps=1000,
                 bound_add=3.5):
    """Exact restatement of rejection_diffeq! (reference src/rejectiondiffeq.jl:76-152 with
    rejectionjump :14-74) for examples/sir-rejection.jl, whose flow is F_dummy: candidate
    times need no ODE, so everything here is exact arithmetic on the uniform stream.

    Returns dict(time, xd[K,4], njumps, nrejected, status, ndraws); status mirrors
    PDMP_ST_* (0 ok / 1 hit n_jumps / 4 bound violated)."""
    r = _Stream(u)
    nu = np.array([[-1, 1, 0, 0], [0, -1, 1, 0], [0, 0, 0, 1]], dtype=np.int64)
    beta, mu = parms
    xd = np.array(xd0, dtype=np.int64)
    bound = beta + bound_add

    def rates():
        S_, I_ = float(xd[0]), float(xd[1])
        return [beta * S_ * I_, mu * I_, beta]

    times, xds = [ti], [xd.copy()]
    tstop = -math.log(r.rand()) / bound + ti
    t, simnj, fict, status = ti, 1, 0, 0
    while t < tf and simnj < n_jumps:
        tc = tstop
        w = rates()
        total = w[0] + w[1] + w[2]
        if not (total < bound):
            status = 4
            break
        reject = r.rand() < 1.0 - total / bound
        dtn = -math.log(r.rand()) / bound
        if tc < tf and not reject:
            ev = pfsample(w, r.rand())
            xd = xd + nu[ev - 1]
            simnj += 1
        else:
            fict += 1
        tstop += dtn
        t = tc
        if t <= tf and not reject:
            times.append(t); xds.append(xd.copy())
    if status == 0:
        if t > tf:
            times.append(tf); xds.append(xd.copy())
        elif t < tf:
            status = 1
    return dict(time=np.array(times), xd=np.array(xds), njumps=simnj, nrejected=fict,
                status=status, ndraws=r.k)
