"""Pins the CPU oracle (oracle/pdmp_oracle.cpp) to the reference's own known-answer tests:
the closed-form samplers of examples/tcp.jl, test/pdmpStiff.jl, examples/pdmpExplosion.jl,
examples/tcp_rejection.jl (committed as tests/golden/closed_forms.npz by make_golden.py),
with the reference's tolerances, and with the north star's tighter ones (rel 1e-8 at ODE
tol 1e-10).  CPU only."""
import numpy as np
import pytest

import closed_forms as cf

ODES = ["tsit5", "dp5"]


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))


@pytest.mark.parametrize("ode", ODES)
def test_tcp_chv_reference_tolerance(orc, P, golden, ode):
    # test/runtests.jl:9-12: norm(res.time - res_a[1], Inf) < 1e-4, nj = 10, default tolerances
    pb = P.models.problem("tcp")
    r = orc.solve(pb, "tcp", algorithm="chv", ode=ode, n_jumps=10, replay=golden["tcp10_u"])
    assert len(r.time) == 10  # S8: n_jumps counts the initial point
    assert np.max(np.abs(r.time - golden["tcp10_t"])) < 1e-4
    assert np.array_equal(r.xd[:, 0], golden["tcp10_xd"])
    assert np.all(r.xd[:, 1] == 1)
    assert r.status[0] == P._abi.ST_HIT_NJUMPS and r.njumps[0] == 10


@pytest.mark.parametrize("ode", ODES)
@pytest.mark.parametrize("tag,nj", [("tcp10", 10), ("tcp700", 700)])
def test_tcp_chv_tight(orc, P, golden, ode, tag, nj):
    pb = P.models.problem("tcp")
    r = orc.solve(pb, "tcp", algorithm="chv", ode=ode, n_jumps=nj, replay=golden[f"{tag}_u"],
                  reltol=1e-10, abstol=1e-18)  # x falls to 1e-9 on this stream: relative control
    assert rel(r.time, golden[f"{tag}_t"]) < 1e-8
    assert np.max(np.abs(r.xc[:, 0] - golden[f"{tag}_x"]) / golden[f"{tag}_x"]) < 1e-8
    assert np.array_equal(r.xd[:, 0], golden[f"{tag}_xd"])


def test_tcp_time_stall_is_a_status(orc, P, golden):
    # On this stream x collapses to 1e-13 and jump times stop advancing in double precision
    # near jump 756: the reference would throw (@assert t < lastjumptime, src/chvdiffeq.jl:145);
    # the ensemble engine returns the partial trajectory with a status instead.
    pb = P.models.problem("tcp")
    r = orc.solve(pb, "tcp", algorithm="chv", ode="tsit5", n_jumps=1000, replay=golden["tcp700_u"],
                  reltol=1e-10, abstol=1e-18)
    assert r.status[0] == P._abi.ST_NO_PROGRESS and 700 < len(r.time) < 1000


@pytest.mark.parametrize("ode", ODES)
def test_stiff_chv(orc, P, golden, ode):
    # test/pdmpStiff.jl:119-127: reltol 1e-8, abstol 1e-11 ; times < 3e-3, final xc < 4e-6
    pb = P.models.problem("pdmp_stiff")
    r = orc.solve(pb, "pdmp_stiff", algorithm="chv", ode=ode, n_jumps=50, replay=golden["stiff_u"],
                  reltol=1e-8, abstol=1e-11)
    assert np.max(np.abs(r.time - golden["stiff_t"])) < 3e-3
    assert abs(r.xc[-1, 0] - golden["stiff_x"][-1]) < 4e-6
    assert np.array_equal(r.xd[:, 0], golden["stiff_xd"])
    r = orc.solve(pb, "pdmp_stiff", algorithm="chv", ode=ode, n_jumps=50, replay=golden["stiff_u"],
                  reltol=1e-10, abstol=1e-12)
    assert rel(r.time, golden["stiff_t"]) < 1e-8
    assert rel(r.xc[:, 0], golden["stiff_x"]) < 1e-7


def test_stiff_with_delta_function(orc, P, golden):
    # test/pdmpStiff.jl:198-215: PDMPProblem(F!, R!, Delta!, 2, ...) — the jumps written as a function
    # instead of nu — gives the same trajectory as the nu form
    pa, pd = P.models.problem("pdmp_stiff"), P.models.problem("pdmp_stiff_delta")
    assert pd.nu.shape == (2, 2) and not pd.nu.any()
    a = orc.solve(pa, "pdmp_stiff", algorithm="chv", ode="tsit5", n_jumps=50, replay=golden["stiff_u"])
    d = orc.solve(pd, "pdmp_stiff_delta", algorithm="chv", ode="tsit5", n_jumps=50, replay=golden["stiff_u"])
    assert np.array_equal(a.time, d.time) and np.array_equal(a.xd, d.xd) and np.array_equal(a.xc, d.xc)
    assert np.max(np.abs(d.time - golden["stiff_t"])) < 1e-3


@pytest.mark.parametrize("ode", ODES)
def test_stiff_chv_tf_limited(orc, P, golden, ode):
    # test/pdmpStiff.jl:142-157: tspan[2] = 4.0: all jump times < tf match (2e-5), last time == tf;
    # pre-jump and post-jump saving give the same final point
    pb = P.models.problem("pdmp_stiff")
    pb.tspan[1] = 4.0
    ana = golden["stiff_t"][golden["stiff_t"] < 4.0]
    r1 = orc.solve(pb, "pdmp_stiff", algorithm="chv", ode=ode, n_jumps=50, replay=golden["stiff_u"])
    r2 = orc.solve(pb, "pdmp_stiff", algorithm="chv", ode=ode, n_jumps=50, replay=golden["stiff_u"],
                   save_positions=(True, False))
    assert np.max(np.abs(r1.time[:-1] - ana)) < 2e-5
    assert r1.time[-1] == 4.0 and r2.time[-1] == 4.0
    assert r1.xc[-1, 0] == r2.xc[-1, 0]
    assert np.array_equal(r1.xd[-1], r2.xd[-1])
    assert r1.status[0] == P._abi.ST_OK
    # pre-jump saves the state just BEFORE each jump: xd lags by one
    assert np.array_equal(r2.xd[1:-1, 0], r1.xd[1:-1, 0] - 1)


@pytest.mark.parametrize("ode", ODES)
def test_explosion_chv(orc, P, golden, ode):
    # test/runtests.jl:14-17: < 1e-4 at default tolerances
    pb = P.models.problem("pdmp_explosion")
    r = orc.solve(pb, "pdmp_explosion", algorithm="chv", ode=ode, n_jumps=50, replay=golden["explosion_u"])
    assert np.max(np.abs(r.time - golden["explosion_t"])) < 1e-4
    assert np.array_equal(r.xd[:, 0], golden["explosion_xd"])


def test_draw_count_pinned(orc, P, golden):
    # examples/pdmpExplosion.jl:76-81: the RNG state after a solve does not depend on the ODE
    # solver => the number of draws is 1 + 2*(n_jumps-1). A stream of exactly that length
    # works; one draw shorter is flagged.
    pb = P.models.problem("pdmp_explosion")
    u = golden["explosion_u"]
    need = 1 + 2 * 49
    r = orc.solve(pb, "pdmp_explosion", algorithm="chv", ode="tsit5", n_jumps=50, replay=u[:need])
    assert r.status[0] == P._abi.ST_HIT_NJUMPS
    r = orc.solve(pb, "pdmp_explosion", algorithm="chv", ode="tsit5", n_jumps=50, replay=u[:need - 1])
    assert r.status[0] == P._abi.ST_REPLAY_EXHAUSTED


@pytest.mark.parametrize("ode", ODES)
def test_tcp_rejection(orc, P, golden, ode):
    # test/runtests.jl:41-44: norm(res.xc[1,1:nj] - res_a[2], Inf) < 1e-5 (default tolerances)
    pb = P.models.problem("tcp_rejection")
    r = orc.solve(pb, "tcp_rejection", algorithm="rejection", ode=ode, n_jumps=50, replay=golden["tcprej_u"])
    assert np.max(np.abs(r.xc[:50, 0] - golden["tcprej_x"])) < 1e-5
    assert np.max(np.abs(r.time[:50] - golden["tcprej_t"])) < 1e-9  # candidate times are exact sums
    assert np.array_equal(r.xd[:50, 0], golden["tcprej_xd"])
    assert r.njumps[0] == 50 and r.nrejected[0] == int(golden["tcprej_nrej"][0])


@pytest.mark.parametrize("ode", ODES)
def test_stiff_rejection(orc, P, golden, ode):
    # test/pdmpStiff.jl:172-179: first 4 jump times < 0.0043
    pb = P.models.problem("pdmp_stiff")
    r = orc.solve(pb, "pdmp_stiff", algorithm="rejection", ode=ode, n_jumps=4, replay=golden["stiffrej_u"])
    assert np.max(np.abs(r.time - golden["stiffrej_t"])) < 0.0043
    assert rel(r.xc[:, 0], golden["stiffrej_x"]) < 1e-5
    assert r.nrejected[0] == int(golden["stiffrej_nrej"][0])


def test_sir_rejection_exact(orc, P, golden):
    # zero flow: everything is exact arithmetic on the uniform stream => bit-exact
    pb = P.models.problem("sir_rejection")
    U = golden["sir_u"]
    r = orc.solve(pb, "sir_rejection", algorithm="rejection", ode="tsit5", trajectories=U.shape[0],
                  n_jumps=1000, replay=U)
    for k in range(U.shape[0]):
        tr = r[k]
        nj, nrej, status, ndraws = golden[f"sir{k}_meta"]
        assert np.array_equal(tr.xd.T, golden[f"sir{k}_xd"])
        assert np.max(np.abs(tr.time - golden[f"sir{k}_t"])) < 1e-12
        assert (tr.njumps, tr.nrejected, tr.status) == (nj, nrej, status)


def test_sir_bound_violation_status(orc, P):
    # SURVEY finding 0.5: the example's bound 3.501 is violable; the reference would @assert
    pb = P.models.problem("sir_rejection")
    r = orc.solve(pb, "sir_rejection", algorithm="rejection", ode="tsit5", trajectories=4000, n_jumps=1000, seed=7)
    h = r.status_histogram()
    assert h.get("BOUND_VIOLATED", 0) > 0 and h.get("OK", 0) > 3800


def test_same_seed_same_trajectory(orc, P):
    # test/runtests.jl:22 ("Call many times the same problem")
    pb = P.models.problem("pdmp_stiff")
    a = orc.solve(pb, "pdmp_stiff", algorithm="chv", ode="tsit5", n_jumps=50, seed=8)
    b = orc.solve(pb, "pdmp_stiff", algorithm="chv", ode="tsit5", n_jumps=50, seed=8)
    c = orc.solve(pb, "pdmp_stiff", algorithm="chv", ode="tsit5", n_jumps=50, seed=9)
    assert np.array_equal(a.time, b.time) and not np.array_equal(a.time, c.time)


def test_sharding_invariance(orc, P):
    # Philox counters use the global trajectory id: a shard reproduces the same trajectories
    pb = P.models.problem("tcp2d")
    full = orc.solve(pb, "tcp2d", algorithm="chv", ode="dp5", trajectories=64, n_jumps=40, seed=3)
    part = orc.solve(pb, "tcp2d", algorithm="chv", ode="dp5", trajectories=32, traj_id_offset=32, n_jumps=40, seed=3)
    a = int(full.offsets[32])
    assert np.array_equal(full.time[a:], part.time) and np.array_equal(full.xd[a:], part.xd)


def test_neuron_eva_delta(orc, P):
    # examples/pdmp_example_eva.jl: Delta resets xc to 0 on reaction 2
    pb = P.models.problem("neuron_eva")
    r = orc.solve(pb, "neuron_eva", algorithm="chv", ode="tsit5", n_jumps=200, seed=1234)
    tr = r[0]
    d = np.diff(tr.xd[0])
    resets = np.where(d == -1)[0] + 1           # reaction 2 fired (xd[1]: 1 -> 0)
    assert len(resets) > 3 and np.all(tr.xc[0, resets] == 0.0)
    assert tr.time[-1] == 100.0                 # test/runtests.jl:53 (time[end] == tf)


def test_last_bit_quirk_flag(orc, P):
    # SURVEY finding 0.6: the reference integrates the last bit at default tolerances
    pb = P.models.problem("pdmp_stiff")
    pb.tspan[1] = 4.0
    a = orc.solve(pb, "pdmp_stiff", algorithm="chv", ode="tsit5", n_jumps=50, seed=8)
    b = orc.solve(pb, "pdmp_stiff", algorithm="chv", ode="tsit5", n_jumps=50, seed=8, ref_quirks=True)
    assert np.array_equal(a.time, b.time) and np.array_equal(a.xc[:-1], b.xc[:-1])
    assert abs(a.xc[-1, 0] - b.xc[-1, 0]) < 1e-2 * max(1.0, abs(a.xc[-1, 0]))


def _rates_cst_closed_form(U, nj):
    # test/testRatesCst.jl: total rate 10 + xd[1] (+ parms[1] = 0) is constant between jumps, so the CHV
    # time change is linear: t_k = sum_j E_j / (10 + j); draws: U0 -> E0, then (U_event, U_E) per jump
    E = -np.log(np.concatenate([[U[0]], U[2::2]]))[:nj - 1]
    return np.concatenate([[0.0], np.cumsum(E / (10.0 + np.arange(nj - 1)))])


def test_rates_cst_closed_form(orc, P):
    U = np.clip(np.random.default_rng(8).random(64), 2.0 ** -53, 1 - 2.0 ** -53)
    pb = P.models.problem("rates_cst")
    r = orc.solve(pb, "rates_cst", algorithm="chv", ode="tsit5", n_jumps=14, replay=U, reltol=1e-10, abstol=1e-12)
    assert np.max(np.abs(r.time - _rates_cst_closed_form(U, 14))) < 1e-10
    assert np.array_equal(r.xd[:, 0], np.arange(14)) and np.all(r.xd[:, 1] == 0)


# ---- RejectionExact (reference src/rejection.jl:3-91, examples/neuron_rejection_exact.jl) -----------

def rejection_exact_python(xc0, xd0, tspan, n_jumps, U):
    """Independent restatement in plain numpy of `solve(problem, Flow)` for the mean-field neuron
    network, written from the Julia (src/rejection.jl:3-91), consuming the uniforms U in order."""
    N = len(xc0)
    f = lambda x: x ** 8
    bound = N * f(1.201)
    it = iter(U)
    x, xd, (t, tf) = np.array(xc0, float), np.array(xd0, np.int64), tspan
    dte = -np.log(next(it))
    time, XC, XD, nsteps, nrej = [t], [x.copy()], [xd.copy()], 1, 0
    while t < tf and nsteps < n_jumps:
        reject = True
        while reject and nsteps < n_jumps:
            t1 = min(tf, t + dte / bound)
            xbar = x.sum() / N
            x = xbar + np.exp(-0.24 * (t1 - t)) * (x - xbar)
            t = t1
            tot = f(x).sum()
            assert tot <= bound
            reject = False if t == tf else (next(it) < 1 - tot / bound)
            dte = -np.log(next(it))
            nrej += reject
        rate = f(x)
        if t < tf:
            thr = next(it) * rate.sum()
            ev = int(min(np.searchsorted(np.cumsum(rate), thr, side="left"), N - 1))   # first i with !(cw_i < thr)
            x = x + 0.98 / N
            x[ev] = 0.0
            xd[ev] += 1
        nsteps += 1
        time.append(t); XC.append(x.copy()); XD.append(xd.copy())
    return np.array(time), np.array(XC), np.array(XD), nsteps, nrej


def test_rejection_exact_oracle_vs_plain_restatement(orc, P):
    pb = P.models.problem("neuron_mf")
    assert pb.nu.shape == (100, 100) and not pb.nu.any()
    U = np.clip(np.random.default_rng(5).random(20000), 2.0 ** -53, 1 - 2.0 ** -53)
    r = orc.solve(pb, "neuron_mf", algorithm="rejection_exact", n_jumps=60, replay=U)
    t, xc, xd, nsteps, nrej = rejection_exact_python(pb.xc0, pb.xd0, pb.tspan, 60, U)
    assert r.njumps[0] == nsteps == 60 and r.nrejected[0] == nrej and len(r.time) == 60
    assert np.array_equal(r.xd, xd) and r.xd.shape == (60, 100)          # all N components are saved (src/rejection.jl:32)
    assert np.max(np.abs(r.time - t)) < 1e-12 and np.max(np.abs(r.xc - xc)) < 1e-12
    # the mean of xc is constant along the flow and moves by J/N - x_ev/N at a jump; a spike resets its neuron
    k = orc.solve(pb, "neuron_mf", algorithm="rejection_exact", n_jumps=60, replay=U, save_c_count=2, save_d_count=3)
    assert k.xc.shape == (60, 2) and k.xd.shape == (60, 3) and np.array_equal(k.xd, xd[:, :3])
    # a horizon shorter than the jumps: the last point sits at tf, with no jump there (src/rejection.jl:64-65,78)
    pb.tspan[1] = 1.0
    s = orc.solve(pb, "neuron_mf", algorithm="rejection_exact", n_jumps=60, replay=U)
    t2, _, xd2, n2, nrej2 = rejection_exact_python(pb.xc0, pb.xd0, pb.tspan, 60, U)
    assert s.time[-1] == 1.0 and s.njumps[0] == n2 and s.nrejected[0] == nrej2 and np.array_equal(s.xd, xd2)
    assert s.status[0] == P._abi.ST_OK and r.status[0] == P._abi.ST_HIT_NJUMPS
