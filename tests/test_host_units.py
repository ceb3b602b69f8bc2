"""CPU unit tests: Philox known answers, pfsample semantics, tableau order conditions,
host mirror of the reference API."""
import numpy as np
import pytest

import closed_forms as cf


def test_philox_known_answers(orc):
    # Random123 kat_vectors, philox4x32-10
    assert orc.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_uniform_stream(orc):
    u = orc.uniforms(1234, 5, 100001)
    assert np.all((u > 0) & (u < 1))
    assert abs(u.mean() - 0.5) < 5e-3 and abs(u.var() - 1 / 12) < 2e-3
    # draw k comes from block k//2: prefix stability
    assert np.array_equal(orc.uniforms(1234, 5, 11), u[:11])
    assert not np.array_equal(orc.uniforms(1234, 6, 11), u[:11])


def test_pfsample_semantics(orc):
    # reference src/utils.jl:46-57: strict `<`, last index absorbs round-off, 1-based
    import ctypes as C
    def pf(w, U):
        a = np.array(w, dtype=np.float64)
        return orc.lib().pdmp_oracle_pfsample(a.ctypes.data_as(C.POINTER(C.c_double)), len(w), U)
    assert pf([1.0, 0.0], 0.999999) == 1            # TCP: rate[2] = 0 => always 1
    assert pf([1.0, 1.0], 0.5) == 1                 # cw == thr stops (strict <)
    assert pf([1.0, 1.0], 0.5000001) == 2
    assert pf([0.0, 0.0, 0.0], 0.3) == 1
    assert pf([1.0, 2.0, 3.0], 1.0) == 3
    rng = np.random.default_rng(0)
    for _ in range(200):
        w = rng.random(4) * (rng.random(4) > 0.3)
        U = rng.random()
        assert pf(w, U) == cf.pfsample(list(w), U)


@pytest.mark.parametrize("ode", ["tsit5", "dp5"])
def test_tableau_order_conditions(orc, ode):
    c, a, bt = orc.tableau(ode)
    A = np.zeros((7, 7))
    A[1:, :6] = a
    b = A[6].copy()
    assert np.allclose(A.sum(1), c, atol=1e-15)
    assert abs(bt.sum()) < 1e-15
    bh = b - bt  # embedded weights
    # order conditions up to 5 for b (17 of them), up to 3 (+) for bhat
    Ac, c2, c3 = A @ c, c ** 2, c ** 3
    conds5 = [
        (b.sum(), 1), (b @ c, 1 / 2), (b @ c2, 1 / 3), (b @ Ac, 1 / 6), (b @ c3, 1 / 4), (b @ (c * Ac), 1 / 8),
        (b @ (A @ c2), 1 / 12), (b @ (A @ Ac), 1 / 24), (b @ c ** 4, 1 / 5), (b @ (c2 * Ac), 1 / 10),
        (b @ (Ac * Ac), 1 / 20), (b @ (c * (A @ c2)), 1 / 15), (b @ (c * (A @ Ac)), 1 / 30),
        (b @ (A @ c3), 1 / 20), (b @ (A @ (c * Ac)), 1 / 40), (b @ (A @ (A @ c2)), 1 / 60),
        (b @ (A @ (A @ Ac)), 1 / 120)]
    for got, want in conds5:
        assert abs(got - want) < 2e-15
    for got, want in [(bh.sum(), 1), (bh @ c, 1 / 2), (bh @ c2, 1 / 3), (bh @ Ac, 1 / 6)]:
        assert abs(got - want) < 1e-14


def test_problem_constructors(P):
    # reference src/problem.jl:168-193
    nu = [[1, 0], [0, -1]]
    a = P.PDMPProblem(None, None, nu, [1.0], [0, 1], [0.0], (0.0, 10.0))
    b = P.PDMPProblem(None, None, lambda *x: None, nu, [1.0], [0, 1], [0.0], (0.0, 10.0))
    c = P.PDMPProblem(None, None, lambda *x: None, 2, [1.0], [0, 1], [0.0], (0.0, 10.0))
    assert a.n_c == 1 and a.n_d == 2 and a.n_r == 2 and b.nu.tolist() == nu
    assert c.nu.shape == (2, 2) and not c.nu.any()
    with pytest.raises(ValueError):
        P.PDMPProblem(None, None, [[1, 0, 0]], [1.0], [0, 1], [0.0], (0.0, 1.0))
    with pytest.raises(TypeError):
        P.PDMPProblem(None, None, nu, [1.0], [0, 1], [0.0])


def test_n_jumps_required_and_shape_checked(orc, P):
    pb = P.models.problem("tcp")
    with pytest.raises(ValueError):
        orc.solve(pb, "tcp", algorithm="chv", n_jumps=None)
    with pytest.raises(ValueError):
        orc.solve(pb, "tcp2d", algorithm="chv", n_jumps=10)   # n_c mismatch
    with pytest.raises(RuntimeError):
        orc.solve(P.models.problem("tcp2d"), "tcp2d", algorithm="rejection", n_jumps=10)  # no bound


def test_result_views(orc, P):
    pb = P.models.problem("morris_lecar")
    r = orc.solve(pb, "morris_lecar", algorithm="chv", ode="tsit5", trajectories=3, n_jumps=30, seed=123)
    assert len(r) == 3
    tr = r[1]
    assert tr.xc.shape == (1, 30) and tr.xd.shape == (4, 30) and tr.time.shape == (30,)
    assert tr.time[0] == 0.0 and np.array_equal(tr.xd[:, 0], [25, 0, 60, 0])
    assert np.all(np.diff(tr.time) > 0)
    # channel counts: xd[0]+xd[1] conserved only by reaction 1; nu row 2 is verbatim [1 -1 0 1]
    d = np.diff(tr.xd, axis=1).T
    allowed = {(-1, 1, 0, 0), (1, -1, 0, 1), (0, 0, -1, 1), (0, 0, 1, -1)}
    assert {tuple(x) for x in d} <= allowed


def test_marshal_transport_rate_kind_and_save_counts(P):
    """Host-side decisions that never reach a GPU: the narrow xd transport is chosen only when it is
    provably exact, rate wrappers map to rate_kind, exact-flow options are forwarded."""
    import ctypes as C
    from pdmp_b200 import _marshal, _abi

    def info(n_c, n_d, n_r, n_parms=1, has_delta=0, name=b"m"):
        i = _abi.ModelInfo()
        i.n_c, i.n_d, i.n_r, i.n_parms, i.has_delta, i.name = n_c, n_d, n_r, n_parms, has_delta, name
        return i

    pb = P.models.problem("tcp", tspan=(0.0, 200.0))
    kw = dict(algorithm="chv", trajectories=4)
    _, o, _ = _marshal.build(pb, info(1, 2, 2), n_jumps=1000, **kw)
    assert o.xd_bytes == 8                                           # default: the reference's Int64 layout
    _, o, keep = _marshal.build(pb, info(1, 2, 2), n_jumps=1000, xd_transport="auto", **kw)
    assert o.xd_bytes == 1                                           # event transport: no Delta, |xd| <= 1 + 999 < 2^23
    assert any(isinstance(k, dict) and "event_nu" in k for k in keep)   # the decoding tables travel with the call
    _, o, _ = _marshal.build(pb, info(1, 2, 2), n_jumps=1000, xd_transport="auto", ind_save_d=[0], **kw)
    assert o.xd_bytes == 2                                           # a component mask: int16 (|xd| <= 1 + 999)
    _, o, _ = _marshal.build(pb, info(1, 2, 2), n_jumps=40000, xd_transport="auto", ind_save_d=[0], **kw)
    assert o.xd_bytes == 4                                           # 40000 jumps could overflow int16
    _, o, _ = _marshal.build(pb, info(1, 2, 2), n_jumps=2 ** 24, xd_transport="auto", no_records=True, **kw)
    assert o.xd_bytes == 4                                           # beyond 2^23: no event transport either
    with pytest.raises(ValueError):
        _marshal.build(pb, info(1, 2, 2, has_delta=1), n_jumps=10, xd_transport="event", **kw)
    _, o, _ = _marshal.build(pb, info(1, 2, 2, has_delta=1), n_jumps=10, xd_transport="auto", **kw)
    assert o.xd_bytes == 4                                           # a Delta may move xd arbitrarily
    with pytest.raises(ValueError):
        _marshal.build(pb, info(1, 2, 2), n_jumps=40000, xd_transport="int16", **kw)
    ex = P.models.EXAMPLES["tcp"]
    for wrap, kind in ((P.VariableRate, 0), (P.ConstantRate, 1)):
        q = P.PDMPProblem("tcp", wrap("tcp"), np.array(ex["nu"]), ex["xc0"], ex["xd0"], ex["parms"], (0.0, 1.0))
        d, _, _ = _marshal.build(q, info(1, 2, 2), n_jumps=5, **kw)
        assert d.rate_kind == kind
    q = P.PDMPProblem("tcp", P.CompositeRate("tcp", "tcp"), np.array(ex["nu"]), ex["xc0"], ex["xd0"], ex["parms"], (0.0, 1.0))
    with pytest.raises(ValueError):                                  # the tcp functor has no constant + variable split
        _marshal.build(q, info(1, 2, 2), n_jumps=5, **kw)
    ic = info(1, 2, 2); ic.composite = 1
    d, _, _ = _marshal.build(q, ic, n_jumps=5, **kw)
    assert d.rate_kind == 2                                          # CompositeRate (src/rate.jl:42-69)
    # ABI v8 options: device list, save_rate, ind_save_c / ind_save_d as bit masks
    _, o, keep = _marshal.build(pb, info(1, 2, 2), n_jumps=5, devices=[1, 0, 1], save_rate=True, ind_save_d=[1], **kw)
    assert (o.n_devices, o.device, o.save_rate, o.save_c_mask, o.save_d_mask) == (3, 1, 1, 0, 0b10)
    assert [o.devices[i] for i in range(3)] == [1, 0, 1]
    with pytest.raises(ValueError):
        _marshal.build(pb, info(1, 2, 2), n_jumps=5, ind_save_c=[1], **kw)   # only component 0 exists
    mf = P.models.problem("neuron_mf")
    d, o, _ = _marshal.build(mf, info(100, 100, 100), algorithm="rejection_exact", trajectories=2, n_jumps=50,
                             save_c_count=2, save_d_count=3)
    assert (o.algorithm, o.save_c_count, o.save_d_count) == (_abi.ALG_REJECTION_EXACT, 2, 3)
    assert isinstance(P.RejectionExactGPU("neuron_mf"), P.AbstractRejection)
    # per-trajectory initial conditions and the shape check against the device functor
    d, _, keep = _marshal.build(pb, info(1, 2, 2), n_jumps=5, algorithm="chv", trajectories=3, xc0=np.ones((3, 1)))
    assert d.xc0_per_traj == 1 and d.xd0_per_traj == 0 and d.n_traj == 3
    with pytest.raises(ValueError):
        _marshal.build(pb, info(2, 2, 2), n_jumps=5, **kw)
