"""CPU checks of the rows added in round 2, on the oracle: CompositeRate, save_rate / rate_hist,
ind_save_c / ind_save_d, and the reference-faithful last bit (ref_quirks)."""
import numpy as np
import pytest

import closed_forms as cf
import julia_rng


def _composite_problem(P, tf):
    ex = P.models.EXAMPLES["rates_composite"]
    mk = lambda R: P.PDMPProblem("rates_composite", R, np.array(ex["nu"]), ex["xc0"], ex["xd0"], ex["parms"], (0.0, tf))
    return mk("rates_composite"), mk(P.VariableRate("rates_composite")), mk(P.CompositeRate("Rcst!", "Rvar!"))


@pytest.mark.parametrize("tf", [100000.0, 0.6])
def test_composite_rate_times_identical(orc, P, tf):
    # test/testRatesComposite.jl:89-105 (tf = 1e5) and :117-131 (tf = 0.6): Random.seed!(8); nj = 20;
    # resvar.time == res0.time and rescmp.time == res0.time
    p0, pvar, pcmp = _composite_problem(P, tf)
    u = julia_rng.stream(8, 64)[None, :]
    r0, rv, rc = (orc.solve(pb, "rates_composite", algorithm="chv", ode="tsit5", n_jumps=20, replay=u) for pb in (p0, pvar, pcmp))
    assert np.array_equal(rv.time, r0.time) and np.array_equal(rc.time, r0.time)
    assert np.array_equal(rc.xc, r0.xc) and np.array_equal(rc.xd, r0.xd)
    assert len(r0.time) == (20 if tf > 1 else len(r0.time)) and r0.time[-1] <= tf
    if tf < 1:
        assert r0.time[-1] == tf and r0.status[0] == P._abi.ST_OK


def test_composite_needs_a_registered_split(orc, P):
    ex = P.models.EXAMPLES["tcp"]
    q = P.PDMPProblem("tcp", P.CompositeRate("a", "b"), np.array(ex["nu"]), ex["xc0"], ex["xd0"], ex["parms"], (0.0, 1.0))
    with pytest.raises(ValueError):
        orc.solve(q, "tcp", algorithm="chv", n_jumps=5)


def test_save_rate_is_the_total_rate_at_each_jump(orc, P):
    # src/chvdiffeq.jl:54: push!(prob.rate_hist, sum(rate)) at every executed jump; for the model of
    # test/testRatesCst.jl the total rate before jump k (k = 1, 2, ...) is 10 + (k - 1) + parms[1]
    pb = P.models.problem("rates_cst")
    u = julia_rng.stream(8, 64)[None, :]
    r = orc.solve(pb, "rates_cst", algorithm="chv", ode="tsit5", n_jumps=14, replay=u, save_rate=True)
    tr = r[0]
    assert len(tr.time) == 14 and np.isnan(r.rate[0])
    assert np.array_equal(r.rate[1:], 10.0 + np.arange(13))
    assert len(tr.rates) == 14                       # the reference's rate_hist: slot 1 (uninitialised there) + one per jump
    # tf-limited: the final point at tf carries no rate and PDMPResult.rates is one shorter than time
    pb.tspan[1] = 0.3
    r = orc.solve(pb, "rates_cst", algorithm="chv", ode="tsit5", n_jumps=14, replay=u, save_rate=True)
    tr = r[0]
    assert tr.time[-1] == 0.3 and np.isnan(r.rate[-1]) and len(tr.rates) == len(tr.time) - 1
    # Rejection: rate_hist gets sum(rate) of the pre-jump state of every ACCEPTED jump (src/rejectiondiffeq.jl:134)
    pb = P.models.problem("sir_rejection")
    r = orc.solve(pb, "sir_rejection", algorithm="rejection", ode="tsit5", n_jumps=1000, replay=julia_rng.stream(1234, 4096)[None, :],
                  save_rate=True)
    xd = r.xd.astype(float)
    want = 0.001 * xd[:-2, 0] * xd[:-2, 1] + 0.01 * xd[:-2, 1] + 0.001      # rates at the state BEFORE each jump
    assert np.allclose(r.rate[1:-1], want, rtol=1e-15, atol=0) and np.isnan(r.rate[[0, -1]]).all()


def test_ind_save_masks(orc, P):
    # src/chv.jl:54-62: ind_save_c / ind_save_d select the saved components
    pb = P.models.problem("morris_lecar", tspan=(0.0, 20.0))
    kw = dict(algorithm="chv", ode="tsit5", trajectories=5, n_jumps=50, seed=3)
    full = orc.solve(pb, "morris_lecar", **kw)
    sub = orc.solve(pb, "morris_lecar", ind_save_d=[1, 3], **kw)
    assert sub.n_d == 2 and sub.xd.shape[1] == 2 and np.array_equal(sub.xd, full.xd[:, [1, 3]])
    assert np.array_equal(sub.time, full.time) and np.array_equal(sub.xc, full.xc)
    assert sub[2].xd.shape == (2, len(sub[2].time))
    with pytest.raises(ValueError):
        orc.solve(pb, "morris_lecar", ind_save_d=[4], **kw)


def test_ref_quirks_last_bit(orc, P, golden):
    """The reference's last bit is a fresh solve at OrdinaryDiffEq's default tolerances 1e-3 / 1e-6
    (src/chvdiffeq.jl:160-168): visible on a model with a non-trivial F, invisible on TCP (F = +-1)."""
    pb = P.models.problem("pdmp_stiff")
    pb.tspan[1] = 4.0
    kw = dict(algorithm="chv", ode="tsit5", n_jumps=50, replay=golden["stiff_u"])
    a = orc.solve(pb, "pdmp_stiff", **kw)
    q = orc.solve(pb, "pdmp_stiff", ref_quirks=True, **kw)
    assert np.array_equal(a.time, q.time) and np.array_equal(a.xc[:-1], q.xc[:-1])
    rel = abs(a.xc[-1, 0] - q.xc[-1, 0]) / abs(a.xc[-1, 0])
    assert 1e-9 < rel < 1e-2            # a looser solve of the same flow: different, but only at the 1e-3 level
    pt = P.models.problem("tcp", tspan=(0.0, 3.0))
    kw = dict(algorithm="chv", ode="tsit5", n_jumps=50, replay=golden["tcp10_u"])
    a, q = orc.solve(pt, "tcp", **kw), orc.solve(pt, "tcp", ref_quirks=True, **kw)
    assert np.allclose(a.xc, q.xc, rtol=1e-12, atol=0)
