"""Loader of the product library libpdmp_b200.so (built in-tree by build.py).

There is NO fallback: if the library is missing or no GPU is usable, calls raise.
"""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# PDMP_B200_LIB selects another build of the same library (kernel A/B experiments: scripts/build_variant.py)
LIB_PATH = os.environ.get("PDMP_B200_LIB") or os.path.join(_HERE, "libpdmp_b200.so")
_LIB = None


class PDMPError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpdmp_b200 error {code}: {msg}")
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.pdmp_last_error.restype = C.c_char_p
        P = C.POINTER
        L.pdmp_model_info_by_id.argtypes = [C.c_int32, P(_abi.ModelInfo)]
        L.pdmp_model_lookup.argtypes = [C.c_char_p, P(_abi.ModelInfo)]
        L.pdmp_solve_ensemble.argtypes = [P(_abi.ProblemDesc), P(_abi.SolveOpts), P(_abi.Result)]
        L.pdmp_result_free.argtypes = [P(_abi.Result)]
        L.pdmp_result_free.restype = None
        L.pdmp_ensemble_create.argtypes = [P(_abi.ProblemDesc), P(_abi.SolveOpts), P(C.c_void_p)]
        L.pdmp_ensemble_upload_ics.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.pdmp_ensemble_run.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.pdmp_ensemble_fetch.argtypes = [C.c_void_p, C.c_int32, P(_abi.Result)]
        L.pdmp_ensemble_run_host.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                             C.c_int32, P(_abi.Result)]
        L.pdmp_ensemble_stats_device.argtypes = [C.c_void_p, P(C.c_void_p), P(C.c_uint64), P(C.c_void_p),
                                                 P(C.c_uint64)]
        L.pdmp_ensemble_counters.argtypes = [C.c_void_p, P(_abi.Result)]
        L.pdmp_ensemble_destroy.argtypes = [C.c_void_p]
        L.pdmp_ensemble_destroy.restype = None
        L.pdmp_fp64_peak_tflops.argtypes = [C.c_int32, P(C.c_double)]
        v = L.pdmp_abi_version()
        if v != _abi.ABI_VERSION:
            raise ImportError(f"libpdmp_b200.so ABI {v} != host mirror ABI {_abi.ABI_VERSION}: rebuild")
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise PDMPError(rc, lib().pdmp_last_error().decode(errors="replace"))


def model_info(name):
    info = _abi.ModelInfo()
    check(lib().pdmp_model_lookup(str(name).encode(), C.byref(info)))
    return info


def list_models():
    L = lib()
    out = []
    for i in range(L.pdmp_model_count()):
        info = _abi.ModelInfo()
        check(L.pdmp_model_info_by_id(i, C.byref(info)))
        out.append(info)
    return out


def device_count():
    return int(lib().pdmp_device_count())


def fp64_peak_tflops(device=0):
    v = C.c_double()
    check(lib().pdmp_fp64_peak_tflops(device, C.byref(v)))
    return v.value
