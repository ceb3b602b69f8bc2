// exact_kernel.cuh — ensemble RejectionExact: thinning with an EXACT flow instead of an ODE solve.
//
// Replaces, for N_traj trajectories at once, the reference's
//     solve(problem::PDMPProblem, Flow::Function; ...)      src/rejection.jl:3-91
//     solve(problem, ::RejectionExact; ...)                 src/rejection.jl:113-116
// for models with MANY continuous variables (examples/neuron_rejection_exact.jl: N = 100 neurons,
// one rate and one discrete counter per neuron).  Such a state does not fit one thread's
// registers, and there is no ODE to integrate, so the parallel shape differs from
// ensemble_kernel.cuh: a GROUP OF G LANES PER TRAJECTORY (G = 8: four trajectories per warp).  Lane l
// of a group holds components PER*l .. PER*l+PER-1 (blocked, so that prefix sums run in index order)
// in registers; totals, the empirical mean and the categorical draw over N rates are group
// reductions / a group prefix scan (shuffles); the uniforms are computed redundantly by every lane
// of the group (no shuffle needed).  Groups pull trajectories from the global cursor.
//
// An exact-flow model provides
//     static constexpr int N, NP;  static constexpr const char* NAME;
//     flow(x, xd, p, t0, t1, W)      x(t0) -> x(t1) in place      Phi(out, xc, xd, parms, [t0, t1])
//     rate(x_i, xd_i, p, t)          rate of reaction i           R(rate, xc, xd, parms, t, false)
//     bound(p)                       constant bound on the total   R(...)[2]
//     delta(x, xd, p, t, ev, W)      the jump                      Delta(xc, xd, parms, t, ind_reaction)
// where W gives the lane's view: W.idx(j) = global index of the lane's j-th component, W.valid(j),
// W.sum(v) = group sum of a lane-local value.
#pragma once
#include "ensemble_kernel.cuh"

namespace pdmp {

// A trajectory is held by a GROUP of G lanes (G = 8, 16 or 32; 32 / G trajectories per warp): the
// scalar part of a candidate (one Philox block, a log, an exp, two divisions: ~2/3 of its
// instructions) is the same for every lane of a group, so narrow groups amortise it over several
// trajectories per warp instruction.  Every lane of a group carries the group's scalars redundantly.
template <int N_, int G_>
struct GroupView {
    static constexpr int N = N_, G = G_, PER = (N_ + G_ - 1) / G_;
    unsigned gl;      // lane within the group
    unsigned gmask;   // the group's lanes: shuffles stay legal when groups diverge from each other
    PDMP_DEV int idx(int j) const { return (int)gl * PER + j; }
    PDMP_DEV bool valid(int j) const { return idx(j) < N; }
    PDMP_DEV double sum(double v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
        return v;
    }
    PDMP_DEV int min_(int v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(gmask, v, o));
        return v;
    }
    PDMP_DEV double scan_incl(double v) const {   // inclusive prefix sum over the group's lanes
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
            const double u = __shfl_up_sync(gmask, v, o, G);
            if ((int)gl >= o) v += u;
        }
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

// The reference's two nested loops (src/rejection.jl:49-52) are flattened: one loop iteration = one
// candidate for every live group; the accepted branch (jump, save, nsteps += 1: 1 candidate in ~35
// for the neuron example) and the (re)initialisation are predicated rare events.
template <class M, int G>
__global__ void __launch_bounds__(128) exact_kernel(const __grid_constant__ KParams P, const ExactOut O) {
    using W = GroupView<M::N, G>;
    constexpr int PER = W::PER, N = M::N;
    const unsigned lane = threadIdx.x & 31u;
    W w{lane % G, (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (lane - lane % G)};
    const int leader = (int)(lane - w.gl);
    unsigned long long c_cands = 0, c_jumps = 0;

    double x[PER];
    int xd[PER];
    DrawStream rng;
    unsigned long long idx = 0;
    double t = 0.0, dte = 0.0;
    long long nsteps = 0, nrej = 0;
    bool need_init = true, dead = false, ex = false;
    const double bnd = M::bound(P.parms);                                           // ppf[2]

    auto draw = [&]() {
        return rng.next(P.rng_mode, P.replay_u + (size_t)idx * P.replay_stride, P.replay_stride,
                        P.traj_id_offset + idx, P.keys, ex);
    };
    auto save = [&](unsigned long long k) {          // _push_time! / push!(xc_hist) / push!(xd_hist), :87-89
        if (k >= O.cap) return;
        const size_t at = (size_t)idx * O.cap + k;
        if (w.gl == 0) O.time[at] = t;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = w.idx(j);
            if (w.valid(j) && i < O.kc) O.xc[at * O.kc + i] = x[j];
            if (w.valid(j) && i < O.kd) O.xd[at * O.kd + i] = (long long)xd[j];
        }
    };
    auto finish = [&](int status) {
        if (w.gl == 0) {
            P.nrec[idx] = (unsigned long long)min((unsigned long long)nsteps, O.cap);
            P.njumps[idx] = nsteps;                                                 // PDMPResult(..., nsteps, n_reject), :90
            P.nrejected[idx] = nrej;
            P.status[idx] = (unsigned char)status;
        }
        need_init = true;
    };

    for (;;) {
        __syncwarp();
        if (need_init && !dead) {
            // ---- init!(problem) + the set-up of src/rejection.jl:8-47 ----
            unsigned long long nx = 0;
            if (w.gl == 0) nx = atomicAdd(P.cursor, 1ull);
            idx = __shfl_sync(w.gmask, nx, leader);
            if (idx >= P.n_traj) dead = true;
            else {
#pragma unroll
                for (int j = 0; j < PER; ++j) {
                    const int i = w.valid(j) ? w.idx(j) : 0;
                    x[j] = w.valid(j) ? P.xc0[(P.xc0_per_traj ? (size_t)idx * N : 0) + i] : 0.0;
                    xd[j] = w.valid(j) ? (int)P.xd0[(P.xd0_per_traj ? (size_t)idx * N : 0) + i] : 0;
                }
                rng.init(); ex = false;
                t = P.t0; nsteps = 1; nrej = 0;
                dte = -log(draw());                                                 // tstop_extended, src/problem.jl:153
                save(0);
                need_init = false;
                if (!(t < P.tf && nsteps < P.n_jumps)) finish(t < P.tf ? PDMP_ST_HIT_NJUMPS : PDMP_ST_OK);   // :49
            }
        }
        if (__all_sync(0xffffffffu, dead)) break;
        if (dead || need_init) continue;

        // ---- one candidate, src/rejection.jl:53-72 ----
        const double t1 = fmin(P.tf, t + dte / bnd);
        M::flow(x, xd, P.parms, t, t1, w);
        t = t1;
        double loc = 0.0;
#pragma unroll
        for (int j = 0; j < PER; ++j) loc += w.valid(j) ? M::rate(x[j], xd[j], P.parms, t) : 0.0;
        const double tot = w.sum(loc);                                              // ppf[1]
        ++c_cands;
        if (!(tot <= bnd)) { finish(PDMP_ST_BOUND_VIOLATED); continue; }            // @assert :63
        bool reject;
        if (t == P.tf) { reject = false; dte = -log(draw()); }                      // :64-65, :69
        else {
            double u_acc, u_exp;   // rand() of :67 then rand() of :69: one Philox call whatever the parity
            rng.next2(P.rng_mode, P.replay_u + (size_t)idx * P.replay_stride, P.replay_stride, P.traj_id_offset + idx,
                      P.keys, ex, u_acc, u_exp);
            reject = u_acc < 1.0 - tot / bnd;
            dte = -log(u_exp);
        }
        if (ex) { finish(PDMP_ST_REPLAY_EXHAUSTED); continue; }
        if (reject) { ++nrej; continue; }                                           // :70-72

        // ---- accepted: the jump (:76-84), then nsteps += 1 and the saved point (:86-89) ----
        if (t < P.tf) {
            // pfsample over the N rates (src/utils.jl:46-57): inclusive prefix sums in index order
            const double thr = draw() * tot;
            double run = 0.0, incl[PER];
#pragma unroll
            for (int j = 0; j < PER; ++j) { run += w.valid(j) ? M::rate(x[j], xd[j], P.parms, t) : 0.0; incl[j] = run; }
            const double excl = w.scan_incl(run) - run;
            int mine = N;                              // first index whose inclusive sum is not < thr; N absorbs round-off
#pragma unroll
            for (int j = PER - 1; j >= 0; --j)
                if (w.valid(j) && !(excl + incl[j] < thr)) mine = w.idx(j) + 1;
            M::delta(x, xd, P.parms, t, w.min_(mine), w);                           // affect!, :83
            ++c_jumps;
            if (ex) { finish(PDMP_ST_REPLAY_EXHAUSTED); continue; }
        }
        save((unsigned long long)nsteps);
        ++nsteps;
        if (!(t < P.tf && nsteps < P.n_jumps)) finish(t < P.tf ? PDMP_ST_HIT_NJUMPS : PDMP_ST_OK);   // :49
    }
    if (w.gl == 0) {
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
    int n, n_parms, group;   // group: lanes per trajectory
    const char* name;
    cudaError_t (*launch)(const KParams& P, const ExactOut& O, int grid, cudaStream_t st);
    cudaError_t (*occupancy)(int* blocks_per_sm);
};
template <class M>
cudaError_t exact_launch(const KParams& P, const ExactOut& O, int grid, cudaStream_t st) {
    exact_kernel<M, M::GROUP><<<grid, 128, 0, st>>>(P, O);
    return cudaGetLastError();
}
template <class M>
cudaError_t exact_occupancy(int* out) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, exact_kernel<M, M::GROUP>, 128, 0);
}
template <class M>
ExactEntry make_exact_entry() {
    return ExactEntry{M::N, M::NP, M::GROUP, M::NAME, &exact_launch<M>, &exact_occupancy<M>};
}

// ---- examples/neuron_rejection_exact.jl: mean-field network of N = 100 neurons ---------------------
struct NeuronMfDev {
    static constexpr int N = 100, NP = 1;
#ifndef PDMP_EXACT_GROUP
#define PDMP_EXACT_GROUP 8
#endif
    static constexpr int GROUP = PDMP_EXACT_GROUP;   // lanes per trajectory: 13 neurons per lane, 4 trajectories per warp
    static constexpr const char* NAME = "neuron_mf";
    // Phi: the empirical mean is constant between jumps; x_i relaxes to it at rate lambda = 0.24  (:12-20)
    template <class W>
    PDMP_DEV static void flow(double (&x)[W::PER], const int (&xd)[W::PER], const double* p, double t0, double t1, const W& w) {
        double loc = 0.0;
#pragma unroll
        for (int j = 0; j < W::PER; ++j) loc += w.valid(j) ? x[j] : 0.0;
        const double xbar = w.sum(loc) / (double)N;
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
