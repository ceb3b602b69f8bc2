"""solve(problem, alg; kwargs) — the reference's entry point (src/chvdiffeq.jl:211-237,
src/rejectiondiffeq.jl:155-176) for the GPU ensemble algorithms, plus the resident
`Ensemble` handle for repeated solves with device-resident inputs."""
import ctypes as C

import numpy as np

from . import _abi, _lib, _marshal
from .algorithms import CHVGPU, RejectionGPU, _GPUAlg


def _common_kwargs(alg, kw):
    kw = dict(kw)
    kw.pop("verbose", None)  # accepted for signature parity, no effect
    if alg.algorithm == "rejection":
        kw.setdefault("save_rate", True)    # the reference's default for Rejection(ode), src/rejectiondiffeq.jl:162
    if kw.pop("finalizer", None) is not None:
        raise NotImplementedError("finalizer is a Julia closure hook (src/chvdiffeq.jl:157); not portable to the device")
    kw.setdefault("reltol", 1e-7)   # reference defaults, src/chvdiffeq.jl:216-217
    kw.setdefault("abstol", 1e-9)
    kw.setdefault("save_positions", (False, True))
    return dict(algorithm=alg.algorithm, ode=alg.ode, precision=alg.precision, **kw)


def solve(problem, alg, *, trajectories=1, **kw):
    """Simulate `trajectories` independent copies of `problem` on the GPU.

    Keyword arguments of the reference `solve` keep their meaning (n_jumps, save_positions,
    reltol, abstol); ensemble additions: trajectories, seed, replay, sample_times, hist,
    hist_bins, device, devices (list of CUDA ordinals: sharded inside the library), traj_id_offset,
    xc0/xd0 (per-trajectory initial conditions), max_records, no_records, save_rate, ind_save_c/ind_save_d
    (0-based component indices), ref_quirks.  Returns an EnsembleResult; `res[i]` is trajectory i as a
    PDMPResult laid out like the reference's.
    """
    if not isinstance(alg, _GPUAlg):
        raise TypeError("alg must be CHVGPU(model) or RejectionGPU(model); the reference's CPU algorithms "
                        "CHV(ode)/Rejection(ode) are not part of this engine")
    L = _lib.lib()
    info = _lib.model_info(alg.model)
    args = _common_kwargs(alg, kw)
    d, o, keep = _marshal.build(problem, info, trajectories=trajectories, **args)
    res = _abi.Result()
    rc = L.pdmp_solve_ensemble(C.byref(d), C.byref(o), C.byref(res))
    if rc == _abi.ERR_CAPACITY:
        # the run reports how many records it wanted: retry once with that capacity
        needed = int(res.records_needed)
        L.pdmp_result_free(C.byref(res))
        o.max_records = needed + needed // 16 + 1024
        res = _abi.Result()
        rc = L.pdmp_solve_ensemble(C.byref(d), C.byref(o), C.byref(res))
    try:
        _lib.check(rc)
        return _marshal.to_numpy(res, copy=True, save_positions=args["save_positions"],
                                 sample_times=args.get("sample_times"), event=_event_tables(keep))
    finally:
        L.pdmp_result_free(C.byref(res))


def _event_tables(keep):
    return next((k for k in keep if isinstance(k, dict) and "event_nu" in k), None)


class Ensemble:
    """Resident ensemble: device buffers + pinned host staging persist across runs.

        ens = Ensemble(problem, CHVGPU("tcp"), trajectories=10**7, n_jumps=1000, ...)
        ens.run(seed)                    # inputs already resident in HBM
        res = ens.fetch()                # D2H into pinned staging; zero-copy numpy views
        ens.upload(xc0=host_array)       # new initial conditions from host memory
    """

    def __init__(self, problem, alg, *, trajectories, **kw):
        L = _lib.lib()
        info = _lib.model_info(alg.model)
        self._args = _common_kwargs(alg, kw)
        d, o, keep = _marshal.build(problem, info, trajectories=trajectories, **self._args)
        self._keep = keep
        self._h = C.c_void_p()
        _lib.check(L.pdmp_ensemble_create(C.byref(d), C.byref(o), C.byref(self._h)))
        self.info = info
        self.n_traj = int(trajectories)

    def upload(self, xc0=None, xd0=None):
        L = _lib.lib()
        xc = xd = None
        pc = pd = None
        per_c = per_d = 0
        if xc0 is not None:
            xc = np.ascontiguousarray(xc0, dtype=np.float64)
            per_c = int(xc.ndim == 2)
            pc = xc.ctypes.data
        if xd0 is not None:
            xd = np.ascontiguousarray(xd0, dtype=np.int64)
            per_d = int(xd.ndim == 2)
            pd = xd.ctypes.data
        _lib.check(L.pdmp_ensemble_upload_ics(self._h, pc, per_c, pd, per_d))

    def run(self, seed=0, stream=None):
        """Asynchronous resident run.  `stream`: None = the ensemble's own stream; otherwise a
        cudaStream_t handle as an integer (e.g. torch.cuda.current_stream().cuda_stream) — the run is
        ordered after everything already queued on it, and work queued on it afterwards sees the results.
        Handle 0 (the legacy default stream, what torch's default stream reports) is passed as
        cudaStreamLegacy, because NULL means "own stream" in the C ABI."""
        if stream is not None and int(stream) == 0:
            stream = 1          # cudaStreamLegacy
        _lib.check(_lib.lib().pdmp_ensemble_run(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, stream))

    def run_host(self, seed=0, xc0=None, xd0=None, records=True, copy=False):
        """One host-buffer step: H2D of the initial conditions (host arrays; None keeps the resident
        ones), the ensemble run, and D2H of every chunk's results overlapped with later chunks."""
        xc = xd = None
        pc = pd = None
        per_c = per_d = 0
        if xc0 is not None:
            xc = np.ascontiguousarray(xc0, dtype=np.float64)
            per_c = int(xc.ndim == 2)
            pc = xc.ctypes.data
        if xd0 is not None:
            xd = np.ascontiguousarray(xd0, dtype=np.int64)
            per_d = int(xd.ndim == 2)
            pd = xd.ctypes.data
        res = _abi.Result()
        rc = _lib.lib().pdmp_ensemble_run_host(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, pc, per_c, pd, per_d,
                                               int(bool(records)), C.byref(res))
        _lib.check(rc)
        return _marshal.to_numpy(res, copy=copy, save_positions=self._args["save_positions"],
                                 sample_times=self._args.get("sample_times"), keep=self, records=records,
                                 event=_event_tables(self._keep))

    def fetch(self, records=True, copy=False):
        res = _abi.Result()
        _lib.check(_lib.lib().pdmp_ensemble_fetch(self._h, int(bool(records)), C.byref(res)))
        return _marshal.to_numpy(res, copy=copy, save_positions=self._args["save_positions"],
                                 sample_times=self._args.get("sample_times"), keep=self, records=records,
                                 event=_event_tables(self._keep))

    def counters(self):
        res = _abi.Result()
        _lib.check(_lib.lib().pdmp_ensemble_counters(self._h, C.byref(res)))
        return dict(rk_attempts=int(res.rk_attempts), rk_rejects=int(res.rk_rejects),
                    aux_attempts=int(res.aux_attempts), total_jumps=int(res.total_jumps),
                    total_candidates=int(res.total_candidates), total_records=int(res.total_records),
                    records_needed=int(res.records_needed), kernel_ms=float(res.kernel_ms),
                    compact_ms=float(res.compact_ms))

    def stats_device(self):
        """(ptr_f64, n_f64, ptr_u64, n_u64): device buffers [count|sum|sumsq] and histogram."""
        pf, nf, ph, nh = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
        _lib.check(_lib.lib().pdmp_ensemble_stats_device(self._h, C.byref(pf), C.byref(nf), C.byref(ph), C.byref(nh)))
        return pf.value, int(nf.value), ph.value, int(nh.value)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().pdmp_ensemble_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
