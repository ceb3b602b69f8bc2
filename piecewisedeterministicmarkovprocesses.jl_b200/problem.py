"""Host-side mirror of the reference's problem/result types.

  PDMPProblem   reference src/problem.jl:132-193 (constructors :168-193)
  PDMPResult    reference src/utils.jl:66-79
  EnsembleResult  ensemble analogue: CSR-ragged saved points + per-trajectory counters +
                  on-GPU statistics (the reference has no ensemble API)
"""
from dataclasses import dataclass, field
from typing import Any, Optional, Tuple

import numpy as np

from . import _abi


def _delta_dummy(xc, xd, parms, t, ind_reaction):  # reference src/jumps.jl:4-6
    return None


class PDMPProblem:
    """PDMPProblem(F, R, [Delta], nu, xc0, xd0, parms, tspan)  — reference src/problem.jl:168-193.

    F, R, Delta are kept for API parity (Julia closures cannot cross to the device; the
    device functor is chosen by the algorithm: ``CHVGPU(model)``).  They may be Python
    callables with the reference signatures, model names, or None.

    Constructors mirrored:
      PDMPProblem(F, R, Delta, nu, xc0, xd0, parms, tspan)
      PDMPProblem(F, R, nu, xc0, xd0, parms, tspan)
      PDMPProblem(F, R, Delta, reaction_number:int, xc0, xd0, parms, tspan)  (all-zero nu)
    """

    def __init__(self, F, R, *args):
        if len(args) == 5:
            Delta = _delta_dummy
            nu, xc0, xd0, parms, tspan = args
        elif len(args) == 6:
            Delta, nu, xc0, xd0, parms, tspan = args
        else:
            raise TypeError("PDMPProblem: expected (F,R,nu,xc0,xd0,parms,tspan) or "
                            "(F,R,Delta,nu,xc0,xd0,parms,tspan)")
        self.F, self.R, self.Delta = F, R, Delta
        self.xc0 = np.array(xc0, dtype=np.float64).reshape(-1)
        self.xd0 = np.array(xd0, dtype=np.int64).reshape(-1)
        if isinstance(nu, (int, np.integer)):  # reaction_number form, src/problem.jl:190-193
            nu = np.zeros((int(nu), self.xd0.size), dtype=np.int64)
        self.nu = np.ascontiguousarray(np.array(nu, dtype=np.int64))
        if self.nu.ndim != 2 or self.nu.shape[1] != self.xd0.size:
            raise ValueError("nu must be [n_reactions x length(xd0)]")
        self.parms = np.array(parms, dtype=np.float64).reshape(-1)
        ti, tf = tspan
        self.tspan = [float(ti), float(tf)]  # mutable like the reference (src/problem.jl:135)

    @property
    def n_c(self):
        return self.xc0.size

    @property
    def n_d(self):
        return self.xd0.size

    @property
    def n_r(self):
        return self.nu.shape[0]


@dataclass
class PDMPResult:
    """One trajectory, laid out like the reference's PDMPResult (src/utils.jl:66-79):
    xc is [n_c, K], xd is [n_d, K] (Julia column k = saved point k)."""
    time: np.ndarray
    xc: np.ndarray
    xd: np.ndarray
    rates: np.ndarray
    save_positions: Tuple[bool, bool]
    njumps: int
    nrejected: int
    status: int = 0


@dataclass
class EnsembleResult:
    n_traj: int
    n_c: int
    n_d: int
    offsets: np.ndarray            # [n_traj+1] uint64
    time: np.ndarray               # [total]
    xc: np.ndarray                 # [total, n_c]
    xd: np.ndarray                 # [total, n_d] int64 (int32 / int16 under a narrow xd_transport)
    njumps: np.ndarray             # [n_traj] int64
    nrejected: np.ndarray          # [n_traj] int64
    status: np.ndarray             # [n_traj] uint8
    stat_count: np.ndarray         # [T, C]
    stat_sum: np.ndarray
    stat_sumsq: np.ndarray
    hist: np.ndarray               # [T, H, B] uint64
    counters: dict = field(default_factory=dict)
    save_positions: Tuple[bool, bool] = (False, True)
    sample_times: Optional[np.ndarray] = None
    shard: Optional[Tuple[int, int]] = None   # (global id offset, count) when produced by solve_sharded / a device block
    rate: Optional[np.ndarray] = None         # [total] total rate at the jump behind every saved point (save_rate)
    # batch statistics (include/pdmp_b200.h): batch = global id mod 64, sums shifted by stat_pivot[c]
    stat_pivot: Optional[np.ndarray] = None   # [C]
    batch_count: Optional[np.ndarray] = None  # [64, T, C]
    batch_sum: Optional[np.ndarray] = None    # sum (v - pivot)
    batch_sumsq: Optional[np.ndarray] = None  # sum (v - pivot)^2
    blocks: Optional[list] = None             # per-device EnsembleResults of a multi-device run
    device: int = 0
    _keep: Any = None
    # set on demand (not dataclass fields, so that __getattr__ can build them lazily):
    #   _event    (nu_ext [n_r + 1, n_d], xd0, first global index) under the event transport
    #   _xd_raw   the reaction indices as shipped (uint8 [total])
    #   _xd_full  xd rebuilt from them

    def _set_event_transport(self, nu, xd0, traj_begin):
        """xd arrived as ONE byte per saved point (the reaction index, 0 = none): keep the bytes, rebuild xd lazily."""
        nu_ext = np.vstack([np.zeros((1, nu.shape[1]), np.int64), nu.astype(np.int64)])
        object.__setattr__(self, "_event", (nu_ext, np.asarray(xd0, np.int64), int(traj_begin)))
        object.__setattr__(self, "_xd_raw", self.__dict__["xd"])
        self.__dict__.pop("xd", None)

    def _event_xd0(self, i):
        nu_ext, xd0, begin = self._event
        return xd0[begin + i] if xd0.ndim == 2 else xd0

    _FLAT = ("offsets", "time", "xc", "xd", "njumps", "nrejected", "status", "rate")

    def _defer_flat_arrays(self):
        """Multi-device result: the flat arrays are the concatenation of the blocks' (a host copy of everything): built on
        first access only, one attribute at a time — `res[i]`, the counters and the statistics never need them."""
        for name in self._FLAT:
            self.__dict__.pop(name, None)
        if self.blocks[0].__dict__.get("_event") is not None:      # event transport: the blocks hold the reaction indices
            object.__setattr__(self, "_event", self.blocks[0]._event)
            self.__dict__.pop("_xd_raw", None)
        object.__setattr__(self, "_block_begin", np.cumsum([0] + [b.n_traj for b in self.blocks]))

    def _concat(self, name):
        bl = self.blocks
        if name == "offsets":
            base = np.cumsum([0] + [int(b.offsets[-1]) for b in bl])
            return np.concatenate([b.offsets[:-1] + np.uint64(base[g]) for g, b in enumerate(bl)] + [np.array([base[-1]], np.uint64)])
        if name == "rate" and bl[0].rate is None:
            return None
        if name == "xd" and bl[0].__dict__.get("_event") is not None:
            return np.concatenate([b.xd for b in bl])            # every block rebuilds its own xd from its reaction indices
        return np.concatenate([getattr(b, name) for b in bl])

    def __getattr__(self, name):
        # only reached for attributes that are not set: the deferred flat arrays of a multi-device result, and `xd`
        # under the event transport
        if (name in EnsembleResult._FLAT or name == "_xd_raw") and self.__dict__.get("blocks") and "_block_begin" in self.__dict__:
            v = self._concat(name)
            object.__setattr__(self, name, v)
            return v
        if name in ("_event", "_xd_full") or (name == "_xd_raw" and not self.__dict__.get("blocks")):
            return None
        if name == "xd" and self.__dict__.get("_event") is not None:
            if self._xd_full is None:
                nu_ext, xd0, begin = self._event
                cs = np.cumsum(nu_ext[self._xd_raw], axis=0)                 # prefix sum of nu rows over the whole CSR
                first = self.offsets[:-1].astype(np.int64)
                cnt = np.diff(self.offsets.astype(np.int64))
                base = cs[first] if len(cs) else cs[:0]                     # a trajectory's first point carries event 0
                x0 = xd0[begin:begin + self.n_traj] if xd0.ndim == 2 else np.broadcast_to(xd0, (self.n_traj, xd0.size))
                object.__setattr__(self, "_xd_full", cs - np.repeat(base, cnt, axis=0) + np.repeat(x0, cnt, axis=0))
            return self._xd_full
        raise AttributeError(name)

    def __len__(self):
        return self.n_traj

    def __getitem__(self, i) -> PDMPResult:
        if self.__dict__.get("blocks") and "_block_begin" in self.__dict__:      # multi-device: straight to the block
            i = int(i) + (self.n_traj if i < 0 else 0)
            g = int(np.searchsorted(self._block_begin, i, side="right")) - 1
            return self.blocks[g][i - int(self._block_begin[g])]
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        if self.__dict__.get("_event") is not None and self._xd_full is None:
            # event transport: this trajectory's xd = xd0 + prefix sum of the nu rows of its reaction indices
            xd_i = self._event_xd0(i)[None, :] + np.cumsum(self._event[0][self._xd_raw[a:b]], axis=0)
        else:
            xd_i = self.xd[a:b]
        # PDMPResult.rates (src/utils.jl:69, filled by save_rate: src/chvdiffeq.jl:54, src/rejectiondiffeq.jl:134): one
        # entry per saved jump after the initial slot, which the reference leaves uninitialised (resize!, src/problem.jl:158)
        rates = np.empty(0)
        if self.rate is not None:
            rates = self.rate[a:b]
            if b > a and np.isnan(rates[-1]) and b - a > 1:
                rates = rates[:-1]            # the final point at tf has no rate entry in the reference
        return PDMPResult(self.time[a:b], self.xc[a:b].T, xd_i.T.astype(np.int64, copy=False), rates,
                          self.save_positions, int(self.njumps[i]), int(self.nrejected[i]),
                          int(self.status[i]))

    def _shifted(self):
        """(n, sum (v - p), sum (v - p)^2, p) per (sample time, component): from the batch buffers when present
        (no cancellation for large means), else from the plain totals with p = 0."""
        if self.batch_count is not None and self.batch_count.size:
            return self.batch_count.sum(0), self.batch_sum.sum(0), self.batch_sumsq.sum(0), self.stat_pivot[None, :]
        return self.stat_count, self.stat_sum, self.stat_sumsq, 0.0

    def mean(self):
        n, a, _, p = self._shifted()
        with np.errstate(invalid="ignore", divide="ignore"):
            return p + a / n

    def var(self):
        """Unbiased sample variance per (sample time, component), from the shifted sums."""
        n, a, q, _ = self._shifted()
        with np.errstate(invalid="ignore", divide="ignore"):
            return (q - a * a / n) / (n - 1.0)

    def batch_means(self):
        """[64, T, C] mean of every batch (NaN for an empty batch)."""
        with np.errstate(invalid="ignore", divide="ignore"):
            return self.stat_pivot[None, None, :] + self.batch_sum / self.batch_count

    def batch_vars(self):
        n, a, q = self.batch_count, self.batch_sum, self.batch_sumsq
        with np.errstate(invalid="ignore", divide="ignore"):
            return (q - a * a / n) / (n - 1.0)

    def mean_stderr(self):
        """Standard error of mean() by batch means: std of the 64 batch means / 8.  No distributional assumption
        (the batches are i.i.d. sub-ensembles of equal size up to one trajectory)."""
        return np.nanstd(self.batch_means(), axis=0, ddof=1) / np.sqrt(np.sum(self.batch_count > 0, axis=0))

    def var_stderr(self):
        """Standard error of var() by batch variances (replaces a Gaussian-kurtosis guess)."""
        return np.nanstd(self.batch_vars(), axis=0, ddof=1) / np.sqrt(np.sum(self.batch_count > 1, axis=0))

    def status_histogram(self):
        if self.__dict__.get("blocks") and "status" not in self.__dict__:
            c = sum(np.bincount(b.status, minlength=len(_abi.STATUS_NAMES)) for b in self.blocks)
        else:
            c = np.bincount(self.status, minlength=len(_abi.STATUS_NAMES))
        return {name: int(c[k]) for k, name in enumerate(_abi.STATUS_NAMES) if c[k]}
