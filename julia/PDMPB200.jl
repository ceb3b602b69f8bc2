# PDMPB200.jl — Julia host side of the B200 ensemble engine.
#
# Adds two algorithm types under the reference's type tree
# (reference src/PiecewiseDeterministicMarkovProcesses.jl:10-15) and the matching
# `SciMLBase.solve(problem::PDMPProblem, algo; kwargs...)` methods (the extension pattern of
# reference src/chvdiffeq.jl:211 and src/rejectiondiffeq.jl:155).  The compute is in
# libpdmp_b200.so behind the C ABI of include/pdmp_b200.h; this file only marshals.
#
#     using PiecewiseDeterministicMarkovProcesses, PDMPB200
#     prob = PDMPProblem(F_tcp!, R_tcp!, nu, xc0, xd0, parms, (0.0, 200.0))      # unchanged
#     ens  = solve(prob, CHVGPU(:tcp); trajectories = 10^7, n_jumps = 1000, seed = 1234)
#     ens[17].time, ens[17].xc, ens[17].xd        # a PDMPResult, laid out like the reference's
#
# NOT EXECUTED in the authoring container (no Julia runtime there): the struct layouts below
# are kept in lock-step with include/pdmp_b200.h by tests/test_abi_cpu.py, which parses this
# file and compares field order/types with the ctypes mirror.
module PDMPB200

import PiecewiseDeterministicMarkovProcesses as PDMP
import SciMLBase
using Libdl

export CHVGPU, RejectionGPU, RejectionExactGPU, EnsembleResult, device_count, list_models

const ABI_VERSION = 8
const libpdmp = Ref{String}(get(ENV, "PDMP_B200_LIB", "libpdmp_b200.so"))

# ---- algorithm types ---------------------------------------------------------------------
"""
    CHVGPU(model; ode = :dp5)

Ensemble CHV on the GPU; replaces `CHV(Tsit5())` (reference src/chvdiffeq.jl:2-4).  `model`
names the compiled-in CUDA device functor for F / R / Delta (Julia closures cannot cross to the
device).  `ode` is `:dp5` (Dormand-Prince 5(4)) or `:tsit5`.
"""
struct CHVGPU <: PDMP.AbstractCHV
    model::Symbol
    ode::Symbol
    precision::Symbol     # :fp64 (default) or :fp32 (float state and stage arithmetic, double time)
    CHVGPU(model; ode = :dp5, precision = :fp64) = new(Symbol(model), ode, precision)
end

"""
    RejectionGPU(model; ode = :dp5)

Ensemble thinning with the model's rate bound; replaces `Rejection(Tsit5())`
(reference src/rejectiondiffeq.jl:1-3).
"""
struct RejectionGPU <: PDMP.AbstractRejection
    model::Symbol
    ode::Symbol
    precision::Symbol
    RejectionGPU(model; ode = :dp5, precision = :fp64) = new(Symbol(model), ode, precision)
end

"""
    RejectionExactGPU(model)

Ensemble thinning with the model's EXACT flow (no ODE solve); replaces `RejectionExact()`
(reference src/rejection.jl:1, :3-91).  One warp per trajectory: for models with many continuous
variables (examples/neuron_rejection_exact.jl, N = 100).
"""
struct RejectionExactGPU <: PDMP.AbstractRejectionExact
    model::Symbol
    ode::Symbol
    precision::Symbol
    RejectionExactGPU(model) = new(Symbol(model), :dp5, :fp64)
end

algorithm_code(::CHVGPU) = Int32(0)        # PDMP_ALG_CHV
algorithm_code(::RejectionGPU) = Int32(1)  # PDMP_ALG_REJECTION
algorithm_code(::RejectionExactGPU) = Int32(2)  # PDMP_ALG_REJECTION_EXACT
ode_code(a) = a.ode === :tsit5 ? Int32(1) : a.ode === :dp5 ? Int32(0) : error("ode must be :dp5 or :tsit5")
precision_code(a) = a.precision === :fp32 ? Int32(1) : a.precision === :fp64 ? Int32(0) : error("precision must be :fp64 or :fp32")

# ---- C structs (include/pdmp_b200.h) -------------------------------------------------------
struct ModelInfo
    id::Int32
    n_c::Int32
    n_d::Int32
    n_r::Int32
    n_parms::Int32
    has_bound::Int32
    has_delta::Int32
    zero_flow::Int32
    exact_flow::Int32
    composite::Int32
    name::NTuple{32, UInt8}
end

struct ProblemDesc
    n_traj::UInt64
    traj_id_offset::UInt64
    xc0::Ptr{Float64}
    xd0::Ptr{Int64}
    nu::Ptr{Int64}
    parms::Ptr{Float64}
    t0::Float64
    tf::Float64
    model_id::Int32
    xc0_per_traj::Int32
    xd0_per_traj::Int32
    nu_col_major::Int32
    n_parms::Int32
    rate_kind::Int32
end

struct SolveOpts
    reltol::Float64
    abstol::Float64
    n_jumps::Int64
    max_attempts::Int64
    seed::UInt64
    replay_u::Ptr{Float64}
    replay_stride::UInt64
    sample_times::Ptr{Float64}
    hist_comp::Ptr{Int32}
    hist_lo::Ptr{Float64}
    hist_hi::Ptr{Float64}
    max_records::UInt64
    algorithm::Int32
    ode::Int32
    precision::Int32
    rng_mode::Int32
    save_pre::Int32
    save_post::Int32
    no_records::Int32
    n_sample_times::Int32
    hist_n::Int32
    hist_bins::Int32
    device::Int32
    tstop_policy::Int32
    ref_quirks::Int32
    xd_bytes::Int32
    save_c_count::Int32
    save_d_count::Int32
    devices::Ptr{Int32}
    save_c_mask::UInt64
    save_d_mask::UInt64
    n_devices::Int32
    save_rate::Int32
end

mutable struct CResult
    n_traj::UInt64
    total_records::UInt64
    records_needed::UInt64
    offsets::Ptr{UInt64}
    time::Ptr{Float64}
    xc::Ptr{Float64}
    xd::Ptr{Cvoid}
    njumps::Ptr{Int64}
    nrejected::Ptr{Int64}
    status::Ptr{UInt8}
    stat_count::Ptr{Float64}
    stat_sum::Ptr{Float64}
    stat_sumsq::Ptr{Float64}
    hist::Ptr{UInt64}
    rk_attempts::UInt64
    rk_rejects::UInt64
    aux_attempts::UInt64
    total_jumps::UInt64
    total_candidates::UInt64
    kernel_ms::Float64
    compact_ms::Float64
    h2d_ms::Float64
    d2h_ms::Float64
    n_sample_times::Int32
    n_comp::Int32
    hist_n::Int32
    hist_bins::Int32
    n_c::Int32
    n_d::Int32
    xd_bytes::Int32
    reserved0::Int32
    owner::Ptr{Cvoid}
    rate::Ptr{Float64}
    stat_pivot::Ptr{Float64}
    batch_count::Ptr{Float64}
    batch_sum::Ptr{Float64}
    batch_sumsq::Ptr{Float64}
    blocks::Ptr{Cvoid}
    traj_begin::UInt64
    n_blocks::Int32
    device::Int32
    CResult() = new()
end

# ---- thin ccall layer ----------------------------------------------------------------------
last_error() = unsafe_string(ccall((:pdmp_last_error, libpdmp[]), Cstring, ()))
device_count() = Int(ccall((:pdmp_device_count, libpdmp[]), Int32, ()))

function check(rc::Integer)
    rc == 0 || error("libpdmp_b200 error $rc: $(last_error())")
    nothing
end

function model_info(name::Symbol)
    v = ccall((:pdmp_abi_version, libpdmp[]), Int32, ())
    v == ABI_VERSION || error("libpdmp_b200.so ABI $v != PDMPB200.jl ABI $ABI_VERSION")
    info = Ref{ModelInfo}()
    check(ccall((:pdmp_model_lookup, libpdmp[]), Int32, (Cstring, Ref{ModelInfo}), String(name), info))
    info[]
end

function list_models()
    n = ccall((:pdmp_model_count, libpdmp[]), Int32, ())
    map(0:n-1) do i
        info = Ref{ModelInfo}()
        check(ccall((:pdmp_model_info_by_id, libpdmp[]), Int32, (Int32, Ref{ModelInfo}), i, info))
        nm = collect(info[].name)
        Symbol(String(nm[1:findfirst(==(0x00), nm)-1]))
    end
end

# ---- result ----------------------------------------------------------------------------------
"""
Ensemble analogue of `PDMPResult` (reference src/utils.jl:66-79): CSR-ragged saved points.
`ens[i]` returns trajectory `i` as a genuine `PDMP.PDMPResult` (xc is `n_c × K`, xd `n_d × K`,
exactly the reference's column-per-saved-point layout — the C arrays are already in that order).
"""
struct EnsembleResult
    offsets::Vector{UInt64}
    time::Vector{Float64}
    xc::Matrix{Float64}        # n_c × total
    xd::Matrix{Int64}          # n_d × total
    njumps::Vector{Int64}
    nrejected::Vector{Int64}
    status::Vector{UInt8}
    stat_count::Matrix{Float64}   # C × T  (column = sample time)
    stat_sum::Matrix{Float64}
    stat_sumsq::Matrix{Float64}
    hist::Array{UInt64, 3}        # bins × hist_n × T
    counters::NamedTuple
    save_positions::Tuple{Bool, Bool}
    rate::Vector{Float64}         # save_rate: total rate at the jump behind every saved point (NaN: none)
end

Base.length(r::EnsembleResult) = length(r.njumps)
function Base.getindex(r::EnsembleResult, i::Integer)
    a, b = Int(r.offsets[i]) + 1, Int(r.offsets[i+1])
    # PDMPResult.rates: one entry per saved jump after the initial slot (which the reference leaves
    # uninitialised, src/problem.jl:158); the final point at tf has no entry there
    rates = isempty(r.rate) ? Float64[] : r.rate[a:b]
    (length(rates) > 1 && isnan(rates[end])) && pop!(rates)
    PDMP.PDMPResult(r.time[a:b], r.xc[:, a:b], r.xd[:, a:b], rates, r.save_positions,
                    r.njumps[i], r.nrejected[i])
end

_copy(p::Ptr{T}, dims...) where {T} = (p == C_NULL || prod(dims) == 0) ? zeros(T, dims...) :
                                      copy(unsafe_wrap(Array, p, dims))

# saved discrete states arrive as Int64, or narrower under `xd_bytes = 4 | 2` (less PCIe traffic);
# they are widened here to the reference's Int64
function _xd(res, tot::Int)
    nd = Int(res.n_d)
    (res.xd == C_NULL || tot == 0) && return zeros(Int64, nd, tot)
    T = res.xd_bytes == 2 ? Int16 : res.xd_bytes == 4 ? Int32 : Int64
    Matrix{Int64}(unsafe_wrap(Array, Ptr{T}(res.xd), (nd, tot)))
end

# event transport (`xd_bytes = 1`): the wire carries ONE byte per saved point, the index of the reaction that changed
# xd since the trajectory's previous saved point (0: none); xd is rebuilt as xd0 + a prefix sum of nu rows
function _xd_event(res, tot::Int, nu::Matrix{Int64}, xd0::AbstractVecOrMat{Int64})
    nd = size(nu, 2)
    out = Matrix{Int64}(undef, nd, tot)
    (res.xd == C_NULL || tot == 0) && return out
    ev = unsafe_wrap(Array, Ptr{UInt8}(res.xd), tot)
    offs = unsafe_wrap(Array, res.offsets, Int(res.n_traj) + 1)
    x = Vector{Int64}(undef, nd)
    for i in 1:Int(res.n_traj)
        x .= xd0 isa AbstractMatrix ? view(xd0, :, Int(res.traj_begin) + i) : xd0
        for k in Int(offs[i])+1:Int(offs[i+1])
            e = Int(ev[k])
            e > 0 && (x .+= view(nu, e, :))
            out[:, k] .= x
        end
    end
    out
end

# ---- solve ----------------------------------------------------------------------------------
"""
    solve(problem::PDMPProblem, algo::Union{CHVGPU, RejectionGPU}; trajectories, n_jumps, kwargs...)

Keyword arguments of the reference `solve` keep their meaning (`n_jumps`, `save_positions`,
`reltol = 1e-7`, `abstol = 1e-9`; reference src/chvdiffeq.jl:211-219); ensemble additions:
`trajectories`, `seed`, `replay` (draws × trajectories matrix of uniforms, consumed in the
reference's order), `sample_times`, `hist = [(component, lo, hi), ...]`, `hist_bins`, `device`,
`traj_id_offset`, `xc0` / `xd0` (n_c × trajectories per-trajectory initial conditions),
`max_records`, `no_records`, `devices` (vector of CUDA ordinals: the ensemble is sharded over them inside the
library call, statistics summed there), `save_rate`, `ind_save_c` / `ind_save_d` (1-based component indices,
reference src/chv.jl:54-62), `ref_quirks` (the final point at `tf` computed exactly like the reference's last bit).
`problem.caract.F/R` are not called: the device functor named by `algo.model` is.
Per-trajectory failures (the reference's `@assert`s, src/chvdiffeq.jl:145,
src/rejectiondiffeq.jl:33) are returned in `status`, never thrown.
"""
function SciMLBase.solve(problem::PDMP.PDMPProblem, algo::Union{CHVGPU, RejectionGPU, RejectionExactGPU};
                         trajectories::Integer = 1, n_jumps = nothing, save_positions = (false, true),
                         reltol = 1e-7, abstol = 1e-9, seed::Integer = 0, replay = nothing,
                         sample_times = Float64[], hist = Tuple{Int, Float64, Float64}[], hist_bins = 0,
                         device = 0, traj_id_offset = 0, xc0 = nothing, xd0 = nothing, max_records = 0,
                         no_records = false, max_attempts = 0, tstop_policy = 0, xd_bytes = 8,
                         save_c_count = 0, save_d_count = 0, devices = Int[], ind_save_c = nothing, ind_save_d = nothing,
                         ref_quirks = false, verbose = false,
                         save_rate = algo isa RejectionGPU,   # the reference's defaults: src/chvdiffeq.jl:218, src/rejectiondiffeq.jl:162
                         finalizer = nothing, kwargs...)
    (n_jumps === nothing || !isfinite(n_jumps)) &&
        error("n_jumps must be given and finite for an ensemble solve")
    finalizer === nothing || error("finalizer is a Julia closure (src/chvdiffeq.jl:157): not portable to the device")
    info = model_info(algo.model)
    car = problem.caract
    nu = Matrix{Int64}(car.pdmpjump.nu)                        # column-major: nu_col_major = 1
    (size(nu, 1) == info.n_r && size(nu, 2) == info.n_d && length(car.xc0) == info.n_c) ||
        error("problem shape does not match device model $(algo.model)")
    (algo isa RejectionExactGPU) == (info.exact_flow != 0) ||
        error("RejectionExactGPU needs an exact-flow device model (and only it can run one)")
    if xd_bytes == 2 || xd_bytes == 1   # narrow transports must be provably exact: |xd0| + (n_jumps - 1) max|nu|, and no Delta on xd
        bound = maximum(abs, xd0 === nothing ? car.xd0 : xd0; init = 0) + max(Int64(n_jumps) - 1, 0) * maximum(abs, nu; init = 0)
        (bound < (xd_bytes == 2 ? 32768 : 2^23) && info.has_delta == 0) ||
            error("xd_bytes = $xd_bytes is not provably exact for this problem (bound $bound)")
        (xd_bytes == 2 || (save_positions[2] && ind_save_d === nothing)) ||
            error("xd_bytes = 1 (event transport) needs post-jump saving and all of xd saved")
    end
    xc = xc0 === nothing ? Vector{Float64}(car.xc0) : Matrix{Float64}(xc0)
    xd = xd0 === nothing ? Vector{Int64}(car.xd0) : Matrix{Int64}(xd0)
    parms = car.parms isa AbstractVector ? Vector{Float64}(car.parms) : Float64[]
    st = Vector{Float64}(sample_times)
    hc = Int32[h[1] - 1 for h in hist]; hlo = Float64[h[2] for h in hist]; hhi = Float64[h[3] for h in hist]
    ru = replay === nothing ? Float64[] : Matrix{Float64}(replay)
    dv = Vector{Int32}(devices)
    _mask(ind) = ind === nothing ? UInt64(0) : reduce(|, (UInt64(1) << (Int(i) - 1) for i in ind); init = UInt64(0))
    rate_kind = car.R isa PDMP.ConstantRate ? 1 : car.R isa PDMP.CompositeRate ? 2 : 0   # src/rate.jl:17-69

    res = CResult()
    GC.@preserve xc xd nu parms st hc hlo hhi ru dv begin
        desc = ProblemDesc(trajectories, traj_id_offset, pointer(xc), pointer(xd), pointer(nu),
                           isempty(parms) ? C_NULL : pointer(parms), problem.tspan[1], problem.tspan[2],
                           info.id, xc0 === nothing ? 0 : 1, xd0 === nothing ? 0 : 1, 1, length(parms),
                           rate_kind)
        function opts(maxrec)
            SolveOpts(reltol, abstol, Int64(n_jumps), max_attempts, UInt64(seed),
                      isempty(ru) ? C_NULL : pointer(ru), isempty(ru) ? 0 : size(ru, 1),
                      isempty(st) ? C_NULL : pointer(st), isempty(hc) ? C_NULL : pointer(hc),
                      isempty(hlo) ? C_NULL : pointer(hlo), isempty(hhi) ? C_NULL : pointer(hhi), maxrec,
                      algorithm_code(algo), ode_code(algo), precision_code(algo), isempty(ru) ? 0 : 1,
                      save_positions[1], save_positions[2], no_records, length(st), length(hc),
                      isempty(hc) ? 0 : hist_bins, isempty(dv) ? device : dv[1], tstop_policy, ref_quirks, xd_bytes,
                      save_c_count, save_d_count, isempty(dv) ? C_NULL : pointer(dv), _mask(ind_save_c), _mask(ind_save_d),
                      length(dv), save_rate)
        end
        rc = ccall((:pdmp_solve_ensemble, libpdmp[]), Int32, (Ref{ProblemDesc}, Ref{SolveOpts}, Ref{CResult}),
                   desc, opts(max_records), res)
        if rc == 5   # PDMP_ERR_CAPACITY: the run reports what it wanted; retry once
            need = res.records_needed
            ccall((:pdmp_result_free, libpdmp[]), Cvoid, (Ref{CResult},), res)
            rc = ccall((:pdmp_solve_ensemble, libpdmp[]), Int32, (Ref{ProblemDesc}, Ref{SolveOpts}, Ref{CResult}),
                       desc, opts(need + need ÷ 16 + 1024), res)
        end
        check(rc)
    end
    n, tot = Int(res.n_traj), Int(res.total_records)
    T, C, H, B = Int(res.n_sample_times), Int(res.n_comp), Int(res.hist_n), Int(res.hist_bins)
    # a multi-device run returns one block per device shard (pdmp_result.blocks): concatenate, offsets rebased
    blocks = res.n_blocks > 1 ? [unsafe_load(Ptr{CResult}(res.blocks), g) for g in 1:Int(res.n_blocks)] : [res]
    base = cumsum([0; [Int(b.total_records) for b in blocks]])
    offs = vcat([_copy(b.offsets, Int(b.n_traj) + 1)[1:end-1] .+ UInt64(base[g]) for (g, b) in enumerate(blocks)]..., UInt64[base[end]])
    cat1(f) = vcat([f(b) for b in blocks]...)
    cat2(f) = hcat([f(b) for b in blocks]...)
    out = EnsembleResult(offs, cat1(b -> _copy(b.time, Int(b.total_records))),
                         cat2(b -> _copy(b.xc, Int(res.n_c), Int(b.total_records))),
                         cat2(b -> res.xd_bytes == 1 ? _xd_event(b, Int(b.total_records), nu, xd) : _xd(b, Int(b.total_records))),
                         cat1(b -> _copy(b.njumps, Int(b.n_traj))), cat1(b -> _copy(b.nrejected, Int(b.n_traj))),
                         cat1(b -> _copy(b.status, Int(b.n_traj))),
                         _copy(res.stat_count, C, T), _copy(res.stat_sum, C, T), _copy(res.stat_sumsq, C, T),
                         _copy(res.hist, B, H, T),
                         (rk_attempts = res.rk_attempts, rk_rejects = res.rk_rejects, aux_attempts = res.aux_attempts,
                          total_jumps = res.total_jumps, total_candidates = res.total_candidates,
                          kernel_ms = res.kernel_ms, compact_ms = res.compact_ms, h2d_ms = res.h2d_ms, d2h_ms = res.d2h_ms),
                         (Bool(save_positions[1]), Bool(save_positions[2])),
                         save_rate ? cat1(b -> _copy(b.rate, Int(b.total_records))) : Float64[])
    ccall((:pdmp_result_free, libpdmp[]), Cvoid, (Ref{CResult},), res)
    out
end

end # module
