"""Python loader of the CPU oracle (oracle/libpdmp_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package.
"""
import ctypes as C
import importlib.util
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import pdmp_b200  # noqa: E402  (host mirror types + ABI structs only; no CUDA needed)
from pdmp_b200 import _abi, _marshal  # noqa: E402

_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libpdmp_oracle.so")
    src = os.path.join(_HERE, "pdmp_oracle.cpp")
    hdr = os.path.join(_ROOT, "include", "pdmp_b200.h")
    stale = (not os.path.exists(so)) or (
        os.path.exists(src) and max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libpdmp_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.pdmp_oracle_last_error.restype = C.c_char_p
        L.pdmp_oracle_solve_ensemble.argtypes = [C.POINTER(_abi.ProblemDesc), C.POINTER(_abi.SolveOpts),
                                                 C.POINTER(_abi.Result)]
        L.pdmp_oracle_result_free.argtypes = [C.POINTER(_abi.Result)]
        L.pdmp_oracle_model_info_by_id.argtypes = [C.c_int32, C.POINTER(_abi.ModelInfo)]
        L.pdmp_oracle_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_double)]
        L.pdmp_oracle_pfsample.argtypes = [C.POINTER(C.c_double), C.c_int32, C.c_double]
        _LIB = L
    return _LIB


def model_info(name):
    L = lib()
    for i in range(L.pdmp_oracle_model_count()):
        info = _abi.ModelInfo()
        L.pdmp_oracle_model_info_by_id(i, C.byref(info))
        if info.name.decode() == name:
            return info
    raise KeyError(name)


def set_threads(n):
    lib().pdmp_oracle_set_threads(int(n))


def max_threads():
    return int(lib().pdmp_oracle_max_threads())


def solve(problem, model, *, algorithm="chv", ode="tsit5", **kw):
    """CPU oracle twin of pdmp_b200.solve: same keyword arguments, same EnsembleResult."""
    info = model_info(model)
    sp = kw.get("save_positions", (False, True))
    d, o, keep = _marshal.build(problem, info, algorithm=algorithm, ode=ode, **kw)
    res = _abi.Result()
    rc = lib().pdmp_oracle_solve_ensemble(C.byref(d), C.byref(o), C.byref(res))
    if rc != 0:
        raise RuntimeError(f"oracle error {rc}: {lib().pdmp_oracle_last_error().decode()}")
    try:
        return _marshal.to_numpy(res, copy=True, save_positions=sp, sample_times=kw.get("sample_times"))
    finally:
        lib().pdmp_oracle_result_free(C.byref(res))


def uniforms(seed, gid, n):
    out = np.empty(n, dtype=np.float64)
    lib().pdmp_oracle_uniforms(seed, gid, n, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().pdmp_oracle_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def tableau(ode):
    c = np.zeros(7); a = np.zeros(36); bt = np.zeros(7)
    p = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
    lib().pdmp_oracle_tableau({"dp5": 0, "tsit5": 1}[ode], p(c), p(a), p(bt))
    return c, a.reshape(6, 6), bt
