"""CPU-only checks of the product library: it loads, exports every symbol include/pdmp_b200.h
declares, the registry matches the oracle's, and compute entry points refuse to run without
a GPU (there is no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(P):
    from pdmp_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.lib()


def test_exports_match_header(P, lib):
    hdr = open(os.path.join(ROOT, "include", "pdmp_b200.h")).read()
    declared = set(re.findall(r"\b(pdmp_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(P._abi.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.pdmp_abi_version() == P._abi.ABI_VERSION
    assert f"#define PDMP_ABI_VERSION {P._abi.ABI_VERSION}" in hdr


def test_struct_sizes(P):
    # natural C layout, no surprises: sizes are multiples of 8 and match the field sums
    assert C.sizeof(P._abi.ModelInfo) == 10 * 4 + 32
    assert C.sizeof(P._abi.ProblemDesc) == 8 * 8 + 4 * 6
    assert C.sizeof(P._abi.SolveOpts) == 8 * 12 + 4 * 16 + 8 * 3 + 4 * 2
    assert C.sizeof(P._abi.Result) == 8 * 3 + 8 * 11 + 8 * 5 + 8 * 4 + 4 * 8 + 8 + 8 * 7 + 4 * 2


def test_registry_matches_oracle(P, lib, orc):
    got = {m.name.decode(): (m.n_c, m.n_d, m.n_r, m.n_parms, m.has_bound, m.has_delta, m.zero_flow)
           for m in P.list_models()}
    assert set(got) == set(P.models.EXAMPLES)
    for name in got:
        o = orc.model_info(name)
        assert got[name] == (o.n_c, o.n_d, o.n_r, o.n_parms, o.has_bound, o.has_delta, o.zero_flow), name
        pb = P.models.problem(name)
        assert (pb.n_c, pb.n_d, pb.n_r) == got[name][:3]
    with pytest.raises(P.PDMPError):
        P.model_info("no_such_model")


def test_no_cpu_fallback(P, lib):
    if P.device_count() > 0:
        pytest.skip("a GPU is present")
    pb = P.models.problem("tcp")
    with pytest.raises(P.PDMPError) as ei:
        P.solve(pb, P.CHVGPU("tcp"), trajectories=4, n_jumps=5)
    assert ei.value.code == P._abi.ERR_NO_DEVICE
    with pytest.raises(P.PDMPError):
        P.fp64_peak_tflops()


def test_argument_validation(P, lib):
    pb = P.models.problem("tcp")
    with pytest.raises(ValueError):
        P.solve(pb, P.CHVGPU("tcp"), trajectories=4)              # n_jumps missing
    with pytest.raises(TypeError):
        P.solve(pb, "CHV(Tsit5())", trajectories=4, n_jumps=5)
    with pytest.raises(ValueError):
        P.solve(pb, P.CHVGPU("tcp2d"), trajectories=4, n_jumps=5)  # shape mismatch
    with pytest.raises(ValueError):
        P.CHVGPU("tcp", ode="rodas5")
    with pytest.raises(P.PDMPError) as ei:
        P.solve(P.models.problem("tcp2d"), P.RejectionGPU("tcp2d"), trajectories=4, n_jumps=5)
    assert ei.value.code == P._abi.ERR_MODEL                       # no bound => no Rejection


_JL2C = {"Int32": C.c_int32, "Int64": C.c_int64, "UInt64": C.c_uint64, "Float64": C.c_double,
         "Ptr{Float64}": C.POINTER(C.c_double), "Ptr{Int64}": C.POINTER(C.c_int64),
         "Ptr{UInt64}": C.POINTER(C.c_uint64), "Ptr{Int32}": C.POINTER(C.c_int32),
         "Ptr{UInt8}": C.POINTER(C.c_uint8), "Ptr{Cvoid}": C.c_void_p, "NTuple{32, UInt8}": C.c_char * 32}


def test_julia_binding_struct_layouts(P):
    """julia/PDMPB200.jl cannot be executed here (no Julia runtime): keep its C struct mirrors in
    lock-step with include/pdmp_b200.h by comparing them field by field with the ctypes mirror."""
    src = open(os.path.join(ROOT, "julia", "PDMPB200.jl")).read()
    assert f"const ABI_VERSION = {P._abi.ABI_VERSION}" in src
    for jl, ct in [("ModelInfo", P._abi.ModelInfo), ("ProblemDesc", P._abi.ProblemDesc),
                   ("SolveOpts", P._abi.SolveOpts), ("CResult", P._abi.Result)]:
        body = re.search(r"struct %s\n(.*?)\n(?:    \w+\(\) = new\(\)\n)?end" % jl, src, re.S).group(1)
        fields = re.findall(r"^\s+(\w+)::([\w{}, ]+)$", body, re.M)
        assert [f for f, _ in fields] == [f for f, _ in ct._fields_], jl
        for (name, jt), (_, ctype) in zip(fields, ct._fields_):
            want = _JL2C[jt]
            same = want is ctype or want._type_ == ctype._type_ or (want is C.c_void_p and issubclass(ctype, C._Pointer))
            assert C.sizeof(want) == C.sizeof(ctype) and same, (jl, name)
    # every entry point the binding ccalls is declared by the header
    hdr = open(os.path.join(ROOT, "include", "pdmp_b200.h")).read()
    for sym in set(re.findall(r"\(:(pdmp_\w+), libpdmp", src)):
        assert re.search(r"\b%s\s*\(" % sym, hdr), sym
