// exact_kernel.cuh — ensemble RejectionExact: thinning with an EXACT flow instead of an ODE solve.
//
// Replaces, for N_traj trajectories at once, the reference's
//     solve(problem::PDMPProblem, Flow::Function; ...)      src/rejection.jl:3-91
//     solve(problem, ::RejectionExact; ...)                 src/rejection.jl:113-116
// for models with MANY continuous variables (examples/neuron_rejection_exact.jl: N = 100 neurons,
// one rate and one discrete counter per neuron).  Such a state does not fit one thread's
// registers, and there is no ODE to integrate, so the parallel shape differs from
// ensemble_kernel.cuh: ONE WARP PER TRAJECTORY.  Lane l holds components PER*l .. PER*l+PER-1
// (blocked, so that prefix sums run in index order) in registers; totals, the empirical mean and the
// categorical draw over N rates are warp reductions / a warp prefix scan; the uniforms are computed
// redundantly by every lane (a warp instruction costs the same for 1 or 32 lanes, and no shuffle is
// needed).  Warps pull trajectories from the global cursor like the lanes of the other kernel.
//
// An exact-flow model provides
//     static constexpr int N, NP;  static constexpr const char* NAME;
//     flow(x, xd, p, t0, t1, W)      x(t0) -> x(t1) in place      Phi(out, xc, xd, parms, [t0, t1])
//     rate(x_i, xd_i, p, t)          rate of reaction i           R(rate, xc, xd, parms, t, false)
//     bound(p)                       constant bound on the total   R(...)[2]
//     delta(x, xd, p, t, ev, W)      the jump                      Delta(xc, xd, parms, t, ind_reaction)
// where W gives the lane's view: W.idx(j) = global index of the lane's j-th component, W.valid(j),
// W.sum(v) = warp sum of a lane-local value.
#pragma once
#include "ensemble_kernel.cuh"

namespace pdmp {

template <int N_>
struct WarpView {
    static constexpr int N = N_, PER = (N_ + 31) / 32;
    unsigned lane;
    PDMP_DEV int idx(int j) const { return (int)lane * PER + j; }
    PDMP_DEV bool valid(int j) const { return idx(j) < N; }
    PDMP_DEV static double sum(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
};

struct ExactOut {
    double* time;        // [n_traj][cap]
    double* xc;          // [n_traj][cap][kc]
    long long* xd;       // [n_traj][cap][kd]
    unsigned long long cap;
    int kc, kd;
};

template <class M>
__global__ void __launch_bounds__(128) exact_kernel(const __grid_constant__ KParams P, const ExactOut O) {
    using W = WarpView<M::N>;
    constexpr int PER = W::PER, N = M::N;
    W w{threadIdx.x & 31u};
    const unsigned lane = w.lane;
    unsigned long long c_cands = 0, c_jumps = 0;

    for (;;) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(P.cursor, 1ull);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= P.n_traj) break;

        // ---- init!(problem) + the set-up of src/rejection.jl:8-47 ----
        double x[PER];
        int xd[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = w.valid(j) ? w.idx(j) : 0;
            x[j] = w.valid(j) ? P.xc0[(P.xc0_per_traj ? (size_t)idx * N : 0) + i] : 0.0;
            xd[j] = w.valid(j) ? (int)P.xd0[(P.xd0_per_traj ? (size_t)idx * N : 0) + i] : 0;
        }
        DrawStream rng; rng.init();
        bool ex = false;
        auto draw = [&]() {
            return rng.next(P.rng_mode, P.replay_u + (size_t)idx * P.replay_stride, P.replay_stride,
                            P.traj_id_offset + idx, P.seed, ex);
        };
        double t = P.t0;
        double dte = -log(draw());                       // simjptimes.tstop_extended (src/problem.jl:153)
        const double bnd = M::bound(P.parms);            // ppf[2]
        long long nsteps = 1, nrej = 0;
        int status = PDMP_ST_OK;
        const size_t base = (size_t)idx * O.cap;
        auto save = [&](unsigned long long k) {          // _push_time! / push!(xc_hist) / push!(xd_hist)
            if (k >= O.cap) return;
            if (lane == 0) O.time[base + k] = t;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int i = w.idx(j);
                if (w.valid(j) && i < O.kc) O.xc[(base + k) * O.kc + i] = x[j];
                if (w.valid(j) && i < O.kd) O.xd[(base + k) * O.kd + i] = (long long)xd[j];
            }
        };
        save(0);

        while (t < P.tf && nsteps < P.n_jumps && status == PDMP_ST_OK) {          // :49
            bool reject = true;
            double rate[PER], tot = 0.0;
            while (reject && nsteps < P.n_jumps) {                                  // :52
                const double t1 = fmin(P.tf, t + dte / bnd);                        // :53
                M::flow(x, xd, P.parms, t, t1, w);                                  // :54-58
                t = t1;
                double loc = 0.0;
#pragma unroll
                for (int j = 0; j < PER; ++j) {
                    rate[j] = w.valid(j) ? M::rate(x[j], xd[j], P.parms, t) : 0.0;
                    loc += rate[j];
                }
                tot = W::sum(loc);                                                  // ppf[1], :62
                ++c_cands;
                if (!(tot <= bnd)) { status = PDMP_ST_BOUND_VIOLATED; break; }      // @assert :63
                if (t == P.tf) reject = false;                                      // :64-65
                else reject = draw() < 1.0 - tot / bnd;                             // :67
                dte = -log(draw());                                                 // :69
                if (reject) ++nrej;                                                 // :70-72
                if (ex) { status = PDMP_ST_REPLAY_EXHAUSTED; break; }
            }
            if (status != PDMP_ST_OK) break;
            if (t < P.tf && !reject) {                                              // :78
                // pfsample over the N rates (src/utils.jl:46-57): inclusive prefix sums in index order
                const double thr = draw() * tot;
                double run = 0.0, incl[PER];
#pragma unroll
                for (int j = 0; j < PER; ++j) { run += rate[j]; incl[j] = run; }
                double scan = run;                                                  // warp inclusive scan of lane totals
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double v = __shfl_up_sync(0xffffffffu, scan, o);
                    if ((int)lane >= o) scan += v;
                }
                const double excl = scan - run;
                // first index whose inclusive sum is not < thr; the last index absorbs round-off
                int mine = N;                                                       // 1-based candidates, N = fallback
#pragma unroll
                for (int j = PER - 1; j >= 0; --j)
                    if (w.valid(j) && !(excl + incl[j] < thr)) mine = w.idx(j) + 1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
                M::delta(x, xd, P.parms, t, mine, w);                               // affect!, :83
                ++c_jumps;
                if (ex) { status = PDMP_ST_REPLAY_EXHAUSTED; break; }
            }
            save((unsigned long long)nsteps);                                       // :87-89
            ++nsteps;                                                               // :86
        }
        if (lane == 0) {
            if (status == PDMP_ST_OK && t < P.tf) status = PDMP_ST_HIT_NJUMPS;
            if (status != PDMP_ST_BOUND_VIOLATED && status != PDMP_ST_REPLAY_EXHAUSTED && (unsigned long long)nsteps > O.cap)
                status = PDMP_ST_OUTPUT_OVERFLOW;
            P.nrec[idx] = (unsigned long long)min((unsigned long long)nsteps, O.cap);
            P.njumps[idx] = nsteps;                                                 // PDMPResult(..., nsteps, n_reject), :90
            P.nrejected[idx] = nrej;
            P.status[idx] = (unsigned char)status;
        }
    }
    if (lane == 0) {
        if (c_cands) atomicAdd(P.counters + CT_CANDS, c_cands);
        if (c_jumps) atomicAdd(P.counters + CT_JUMPS, c_jumps);
    }
}

// strided per-trajectory buffers -> CSR: one warp per trajectory, coalesced copies
__global__ void __launch_bounds__(128) exact_compact_kernel(const KParams P, const ExactOut O,
                                                            const unsigned long long* __restrict__ offsets,
                                                            double* __restrict__ time, double* __restrict__ xc,
                                                            long long* __restrict__ xd) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long warp0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long i = warp0; i < P.n_traj; i += nwarps) {
        const unsigned long long n = P.nrec[i], off = offsets[i], base = i * O.cap;
        for (unsigned long long k = lane; k < n; k += 32) time[off + k] = O.time[base + k];
        for (unsigned long long k = lane; k < n * O.kc; k += 32) xc[off * O.kc + k] = O.xc[base * O.kc + k];
        for (unsigned long long k = lane; k < n * O.kd; k += 32) xd[off * O.kd + k] = O.xd[base * O.kd + k];
    }
}

struct ExactEntry {
    int n, n_parms;
    const char* name;
    cudaError_t (*launch)(const KParams& P, const ExactOut& O, int grid, cudaStream_t st);
    cudaError_t (*occupancy)(int* blocks_per_sm);
};
template <class M>
cudaError_t exact_launch(const KParams& P, const ExactOut& O, int grid, cudaStream_t st) {
    exact_kernel<M><<<grid, 128, 0, st>>>(P, O);
    return cudaGetLastError();
}
template <class M>
cudaError_t exact_occupancy(int* out) { return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, exact_kernel<M>, 128, 0); }
template <class M>
ExactEntry make_exact_entry() { return ExactEntry{M::N, M::NP, M::NAME, &exact_launch<M>, &exact_occupancy<M>}; }

// ---- examples/neuron_rejection_exact.jl: mean-field network of N = 100 neurons ---------------------
struct NeuronMfDev {
    static constexpr int N = 100, NP = 1;
    static constexpr const char* NAME = "neuron_mf";
    // Phi: the empirical mean is constant between jumps; x_i relaxes to it at rate lambda = 0.24  (:12-20)
    template <class W>
    PDMP_DEV static void flow(double (&x)[W::PER], const int (&xd)[W::PER], const double* p, double t0, double t1, const W& w) {
        double loc = 0.0;
#pragma unroll
        for (int j = 0; j < W::PER; ++j) loc += w.valid(j) ? x[j] : 0.0;
        const double xbar = W::sum(loc) / (double)N;
        const double e = exp(-0.24 * (t1 - t0));
#pragma unroll
        for (int j = 0; j < W::PER; ++j) x[j] = xbar + e * (x[j] - xbar);
    }
    PDMP_DEV static double f(double v) { const double v2 = v * v, v4 = v2 * v2; return v4 * v4; }   // x^8, :8-10
    PDMP_DEV static double rate(double x, int xd, const double* p, double t) { return f(x); }          // :26-28
    PDMP_DEV static double bound(const double* p) { return (double)N * f(1.201); }                     // :23
    // Delta_xc_mf: every neuron receives J/N, the spiking one is reset and counted  (:38-47)
    template <class W>
    PDMP_DEV static void delta(double (&x)[W::PER], int (&xd)[W::PER], const double* p, double t, int ev, const W& w) {
#pragma unroll
        for (int j = 0; j < W::PER; ++j) {
            x[j] += 0.98 / (double)N;
            if (w.idx(j) + 1 == ev) { x[j] = 0.0; xd[j] += 1; }
        }
    }
};

}  // namespace pdmp
