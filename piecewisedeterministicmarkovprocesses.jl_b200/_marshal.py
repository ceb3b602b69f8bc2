"""Marshalling between the Python host mirror and the C ABI structs (include/pdmp_b200.h).
Library-agnostic: the product loader (_lib.py) and the oracle loader (oracle/oracle.py, test
infrastructure) both use it with their own library handle."""
import ctypes as C

import numpy as np

from . import _abi
from .problem import EnsembleResult

_ODE = {"dp5": _abi.ODE_DP5, "tsit5": _abi.ODE_TSIT5, _abi.ODE_DP5: _abi.ODE_DP5,
        _abi.ODE_TSIT5: _abi.ODE_TSIT5}
_ALG = {"chv": _abi.ALG_CHV, "rejection": _abi.ALG_REJECTION, "rejection_exact": _abi.ALG_REJECTION_EXACT}


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype)) if a is not None and a.size else C.POINTER(ctype)()


def build(problem, info, *, algorithm, ode="dp5", trajectories=1, traj_id_offset=0,
          n_jumps=None, save_positions=(False, True), reltol=1e-7, abstol=1e-9, seed=0,
          replay=None, sample_times=None, hist=None, hist_bins=0, precision="fp64",
          device=0, max_attempts=0, max_records=0, no_records=False, tstop_policy=0,
          ref_quirks=False, xc0=None, xd0=None, xd_transport="int64", save_c_count=0, save_d_count=0,
          devices=None, save_rate=False, ind_save_c=None, ind_save_d=None):
    """Returns (desc, opts, keepalive).  `info` is the model's pdmp_model_info.

    xc0 / xd0 override the problem's (broadcast) initial conditions with per-trajectory
    arrays [trajectories, n_c] / [trajectories, n_d].
    hist: list of (component, lo, hi).
    xd_transport: "int64" (reference layout), "int32" (always exact), "int16", "event" (ONE byte per saved
    point: the index of the reaction that produced it; xd is rebuilt on access as xd0 + a prefix sum of nu
    rows — for models without a Delta and |xd| provably below 2^23), or "auto" = the narrowest of these
    that is provably exact (int32 when the model has a Delta, which may move xd arbitrarily).
    devices: list of CUDA ordinals; the ensemble is sharded over them inside the library call.
    save_rate: also return the total rate at every saved jump (reference `save_rate` / PDMPResult.rates).
    ind_save_c / ind_save_d: 0-based indices of the continuous / discrete components to save (the
    reference's 1-based `ind_save_c` / `ind_save_d`, src/chv.jl:54-62); None saves all."""
    if n_jumps is None or not np.isfinite(n_jumps):
        raise ValueError("n_jumps must be given and finite for an ensemble solve (the reference "
                         "default Inf64 has no meaning with heavy-tailed jump counts)")
    if not (problem.tspan[0] < problem.tspan[1]):
        raise ValueError("tspan must satisfy t0 < tf")
    if problem.n_c != info.n_c or problem.n_d != info.n_d or problem.n_r != info.n_r:
        raise ValueError(
            f"problem shape (n_c={problem.n_c}, n_d={problem.n_d}, n_r={problem.n_r}) does not match "
            f"device model '{info.name.decode()}' ({info.n_c}, {info.n_d}, {info.n_r})")
    if problem.parms.size < info.n_parms:
        raise ValueError(f"model '{info.name.decode()}' reads {info.n_parms} parameters, got {problem.parms.size}")
    keep = []
    n = int(trajectories)
    d = _abi.ProblemDesc()
    d.n_traj = n
    d.traj_id_offset = int(traj_id_offset)
    if xc0 is None:
        xc = np.ascontiguousarray(problem.xc0, dtype=np.float64); d.xc0_per_traj = 0
    else:
        xc = np.ascontiguousarray(xc0, dtype=np.float64).reshape(n, info.n_c); d.xc0_per_traj = 1
    if xd0 is None:
        xd = np.ascontiguousarray(problem.xd0, dtype=np.int64); d.xd0_per_traj = 0
    else:
        xd = np.ascontiguousarray(xd0, dtype=np.int64).reshape(n, info.n_d); d.xd0_per_traj = 1
    nu = np.ascontiguousarray(problem.nu, dtype=np.int64)
    parms = np.ascontiguousarray(problem.parms, dtype=np.float64)
    keep += [xc, xd, nu, parms]
    d.xc0, d.xd0, d.nu, d.parms = _ptr(xc, C.c_double), _ptr(xd, C.c_int64), _ptr(nu, C.c_int64), _ptr(parms, C.c_double)
    d.t0, d.tf = problem.tspan
    d.model_id = info.id
    d.nu_col_major = 0
    d.n_parms = parms.size
    kind = getattr(problem.R, "rate_kind", 0)     # rate.py wrappers; a bare R is a VariableRate (src/problem.jl:55)
    if kind == 2 and not info.composite:
        raise ValueError(f"CompositeRate: device model '{info.name.decode()}' is not registered with a constant + variable "
                         "split of its total rate (csrc/models.cuh: COMPOSITE_OK, rtot_cst, rtot_var)")
    d.rate_kind = kind

    o = _abi.SolveOpts()
    o.reltol, o.abstol = float(reltol), float(abstol)
    o.n_jumps = int(n_jumps)
    o.max_attempts = int(max_attempts)
    o.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    if replay is not None:
        ru = np.ascontiguousarray(replay, dtype=np.float64)
        if ru.ndim == 1:
            ru = ru.reshape(1, -1)
        if ru.shape[0] != n:
            raise ValueError("replay must be [trajectories, draws]")
        keep.append(ru)
        o.replay_u = _ptr(ru, C.c_double)
        o.replay_stride = ru.shape[1]
        o.rng_mode = _abi.RNG_REPLAY
    else:
        o.rng_mode = _abi.RNG_PHILOX
    st = np.ascontiguousarray(sample_times if sample_times is not None else [], dtype=np.float64)
    if st.size and (np.any(np.diff(st) <= 0) or st[0] <= d.t0 or st[-1] > d.tf):
        raise ValueError("sample_times must be increasing and within (t0, tf]")
    keep.append(st)
    o.sample_times = _ptr(st, C.c_double)
    o.n_sample_times = st.size
    hist = list(hist or [])
    hc = np.array([h[0] for h in hist], dtype=np.int32)
    hlo = np.array([h[1] for h in hist], dtype=np.float64)
    hhi = np.array([h[2] for h in hist], dtype=np.float64)
    keep += [hc, hlo, hhi]
    o.hist_comp, o.hist_lo, o.hist_hi = _ptr(hc, C.c_int32), _ptr(hlo, C.c_double), _ptr(hhi, C.c_double)
    o.hist_n = len(hist)
    o.hist_bins = int(hist_bins) if hist else 0
    o.max_records = int(max_records)
    o.algorithm = _ALG[algorithm]
    o.ode = _ODE[ode]
    o.precision = {"fp64": _abi.PREC_FP64, "fp32": _abi.PREC_FP32}[precision]
    o.save_pre, o.save_post = int(bool(save_positions[0])), int(bool(save_positions[1]))
    o.no_records = int(bool(no_records))
    o.device = int(device)
    o.tstop_policy = int(tstop_policy)
    o.ref_quirks = int(bool(ref_quirks))
    bound = int(np.max(np.abs(xd), initial=0)) + max(int(n_jumps) - 1, 0) * int(np.max(np.abs(nu), initial=0))
    if xd_transport == "auto":
        xd_transport = "int32" if info.has_delta else ("event" if (bound < 2 ** 23 and ind_save_d is None and save_positions[1]) else
                                                        "int16" if bound < 32768 else "int32")
    if xd_transport == "event" and (bound >= 2 ** 23 or info.has_delta or ind_save_d is not None or not save_positions[1]):
        raise ValueError("xd_transport='event' needs a model without a Delta, post-jump saving, all of xd saved and "
                         "|xd| < 2^23 (bound %d)" % bound)
    if xd_transport == "int16" and (bound >= 32768 or info.has_delta):
        raise ValueError("xd_transport='int16' is not provably exact for this problem (bound %d)" % bound)
    o.xd_bytes = {"int64": 8, "int32": 4, "int16": 2, "event": 1}[xd_transport]
    if xd_transport == "event":     # what the bindings need to rebuild xd from the reaction indices
        keep.append(dict(event_nu=nu.copy(), event_xd0=xd.copy()))
    o.save_c_count, o.save_d_count = int(save_c_count), int(save_d_count)
    if devices is not None:
        dv = np.ascontiguousarray(devices, dtype=np.int32).reshape(-1)
        if dv.size == 0:
            raise ValueError("devices must not be empty")
        keep.append(dv)
        o.devices, o.n_devices = _ptr(dv, C.c_int32), dv.size
        o.device = int(dv[0])
    o.save_rate = int(bool(save_rate))

    def mask(ind, n, what):
        if ind is None:
            return 0
        ind = sorted(set(int(i) for i in ind))
        if not ind or ind[0] < 0 or ind[-1] >= n:
            raise ValueError(f"{what} must be indices in [0, {n})")
        return sum(1 << i for i in ind)
    o.save_c_mask, o.save_d_mask = mask(ind_save_c, info.n_c, "ind_save_c"), mask(ind_save_d, info.n_d, "ind_save_d")
    return d, o, keep


def _arr(ptr, shape, dtype, copy):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype=dtype)
    if isinstance(ptr, int):   # void* field: raw address
        ptr = C.cast(C.c_void_p(ptr), C.POINTER(C.c_uint8))
        a = np.ctypeslib.as_array(ptr, shape=(n * np.dtype(dtype).itemsize,)).view(dtype).reshape(shape)
    else:
        a = np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).reshape(shape)
    return a.copy() if copy else a


def to_numpy(res, *, copy=True, save_positions=(False, True), sample_times=None, keep=None,
             records=True, event=None):
    """event: dict(event_nu, event_xd0) when the run used xd_transport='event' (see build())."""
    if int(res.n_blocks) > 1:
        return _from_blocks(res, copy=copy, save_positions=save_positions, sample_times=sample_times, keep=keep, records=records,
                            event=event)
    n, tot = int(res.n_traj), int(res.total_records) if records else 0
    T, Cc, H, B = res.n_sample_times, res.n_comp, res.hist_n, res.hist_bins
    nc, nd = res.n_c, res.n_d
    out = EnsembleResult(
        n_traj=n, n_c=nc, n_d=nd,
        offsets=_arr(res.offsets, (n + 1,), np.uint64, copy) if records else np.zeros(n + 1, np.uint64),
        time=_arr(res.time, (tot,), np.float64, copy),
        xc=_arr(res.xc, (tot, nc), np.float64, copy),
        xd=(_arr(res.xd, (tot,), np.uint8, copy) if int(res.xd_bytes) == 1 else
            _arr(res.xd, (tot, nd), {8: np.int64, 4: np.int32, 2: np.int16}[int(res.xd_bytes) or 8], copy)),
        njumps=_arr(res.njumps, (n,), np.int64, copy),
        nrejected=_arr(res.nrejected, (n,), np.int64, copy),
        status=_arr(res.status, (n,), np.uint8, copy),
        stat_count=_arr(res.stat_count, (T, Cc), np.float64, copy),
        stat_sum=_arr(res.stat_sum, (T, Cc), np.float64, copy),
        stat_sumsq=_arr(res.stat_sumsq, (T, Cc), np.float64, copy),
        hist=_arr(res.hist, (T, H, B), np.uint64, copy),
        rate=_arr(res.rate, (tot,), np.float64, copy) if res.rate else None,
        stat_pivot=_arr(res.stat_pivot, (Cc,), np.float64, True) if res.batch_count else None,
        batch_count=_arr(res.batch_count, (_abi.STAT_BATCHES, T, Cc), np.float64, copy) if res.batch_count else None,
        batch_sum=_arr(res.batch_sum, (_abi.STAT_BATCHES, T, Cc), np.float64, copy) if res.batch_count else None,
        batch_sumsq=_arr(res.batch_sumsq, (_abi.STAT_BATCHES, T, Cc), np.float64, copy) if res.batch_count else None,
        counters=dict(rk_attempts=int(res.rk_attempts), rk_rejects=int(res.rk_rejects),
                      aux_attempts=int(res.aux_attempts), total_jumps=int(res.total_jumps),
                      total_candidates=int(res.total_candidates), kernel_ms=float(res.kernel_ms),
                      compact_ms=float(res.compact_ms), h2d_ms=float(res.h2d_ms),
                      d2h_ms=float(res.d2h_ms), total_records=int(res.total_records),
                      records_needed=int(res.records_needed)),
        save_positions=tuple(bool(x) for x in save_positions),
        sample_times=None if sample_times is None else np.array(sample_times, dtype=np.float64),
        _keep=keep,
    )
    if int(res.xd_bytes) == 1:
        if event is None:
            raise ValueError("event transport: the decoding tables (nu, xd0) are missing")
        out._set_event_transport(event["event_nu"], event["event_xd0"], int(res.traj_begin))
    return out


def _from_blocks(res, *, copy, save_positions, sample_times, keep, records, event=None):
    """Multi-device result: one block per device shard (include/pdmp_b200.h, pdmp_result.blocks).  The blocks
    stay separate views (`out.blocks`); `res[i]` goes straight to its block; the flat arrays of the EnsembleResult
    (their concatenation, offsets rebased: a host copy of everything) are built on first access only."""
    blocks = [to_numpy(res.blocks[g], copy=copy, save_positions=save_positions, sample_times=sample_times, keep=keep,
                       records=records, event=event) for g in range(int(res.n_blocks))]
    for g, b in enumerate(blocks):
        b.shard = (int(res.blocks[g].traj_begin), b.n_traj)
        b.device = int(res.blocks[g].device)
    T, Cc, H, B = res.n_sample_times, res.n_comp, res.hist_n, res.hist_bins
    empty = np.zeros(0)
    out = EnsembleResult(
        n_traj=int(res.n_traj), n_c=res.n_c, n_d=res.n_d, offsets=empty, time=empty, xc=empty, xd=empty,
        njumps=empty, nrejected=empty, status=empty,
        stat_count=_arr(res.stat_count, (T, Cc), np.float64, True), stat_sum=_arr(res.stat_sum, (T, Cc), np.float64, True),
        stat_sumsq=_arr(res.stat_sumsq, (T, Cc), np.float64, True), hist=_arr(res.hist, (T, H, B), np.uint64, True),
        stat_pivot=_arr(res.stat_pivot, (Cc,), np.float64, True) if res.batch_count else None,
        batch_count=_arr(res.batch_count, (_abi.STAT_BATCHES, T, Cc), np.float64, True) if res.batch_count else None,
        batch_sum=_arr(res.batch_sum, (_abi.STAT_BATCHES, T, Cc), np.float64, True) if res.batch_count else None,
        batch_sumsq=_arr(res.batch_sumsq, (_abi.STAT_BATCHES, T, Cc), np.float64, True) if res.batch_count else None,
        counters=dict(rk_attempts=int(res.rk_attempts), rk_rejects=int(res.rk_rejects), aux_attempts=int(res.aux_attempts),
                      total_jumps=int(res.total_jumps), total_candidates=int(res.total_candidates),
                      kernel_ms=float(res.kernel_ms), compact_ms=float(res.compact_ms), h2d_ms=float(res.h2d_ms),
                      d2h_ms=float(res.d2h_ms), total_records=int(res.total_records), records_needed=int(res.records_needed)),
        save_positions=tuple(bool(x) for x in save_positions),
        sample_times=None if sample_times is None else np.array(sample_times, dtype=np.float64), _keep=keep)
    out.blocks = blocks
    out._defer_flat_arrays()      # offsets / time / xc / xd / ... = concatenation of the blocks, on first access
    return out
