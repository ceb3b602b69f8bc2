// ensemble_kernel.cuh — the persistent ensemble kernel: one PDMP trajectory per thread, all
// state in registers, finished lanes refilled from a global work cursor.
//
// Replaces, for N trajectories at once:
//   CHV        chv_diffeq! + chvjump + (chv::CHV)(xdot,x,caract,t)   reference src/chvdiffeq.jl:6-16,24-74,76-170
//   Rejection  rejection_diffeq! + rejectionjump                     reference src/rejectiondiffeq.jl:5-8,14-74,76-152
//   init!      src/problem.jl:150-161      pfsample  src/utils.jl:46-57      affect!  src/jumps.jl:28-36
//   saving     _push_time!/_push_xc!/_push_xd!  src/problem.jl:146-148
//
// Loop shape (phase scheduler, see the kernel): a warp alternates between three warp-uniform phases —
// RK (an inner loop of adaptive step attempts for the running lanes, one ballot per attempt), JUMP
// (the threshold section for all lanes that landed, preceded by a converged sampling pass) and
// REFILL (finished lanes pull new trajectories from the global cursor with a warp-aggregated atomic).
// The integrator state stays in registers throughout.
#pragma once
#include <algorithm>

#include "device_core.cuh"
#include "../../include/pdmp_b200.h"

namespace pdmp {

constexpr int MAXP = 16, MAXR = 8, MAXD = 8, MAXH = 4;
constexpr int BLOCK = 128;           // threads per CTA
constexpr int REC_BLOCK = 16;        // records per pool block (allocation granule)

// counters[] slots
enum { CT_ATTEMPTS = 0, CT_REJECTS, CT_AUX, CT_JUMPS, CT_CANDS, CT_N };

struct KParams {
    // problem
    unsigned long long n_traj, traj_id_offset;
    const double* xc0;
    const long long* xd0;
    int xc0_per_traj, xd0_per_traj;
    double parms[MAXP];
    int nu[MAXR * MAXD];  // [r * ND + d]
    double t0, tf;
    int rate_const;   // ConstantRate: the CHV field re-uses the total rate of the post-jump state
    int rate_composite;  // CompositeRate: the constant part is cached from jump to jump, the variable part evaluated per stage
    int ref_quirks;   // the final point at tf the way the reference computes it (default tolerances, caract.xc at tprev)
    // options
    double reltol, abstol;
    long long n_jumps, max_attempts;
    int save_pre, save_post, tstop_policy, rng_mode;
    int jump_thresh, refill_thresh;  // phase scheduler: lanes that must wait before JUMP / REFILL run
    int jump_cap_q;                  // ... and the cap on jump_thresh as quarters of the live lanes (2: half)
    unsigned long long seed;
    PhiloxKeys keys;                 // round keys of `seed` (philox_round_keys)
    const double* replay_u;
    unsigned long long replay_stride;
    // sampling: the ensemble kernel only RECORDS the state at the fixed sample times, one row of C = NC + ND
    // doubles per (trajectory, sample time); stats_reduce_kernel (stats_kernel.cuh) turns the rows into batch
    // moments and histograms in a fixed order (reproducible sums, no atomics on the integration path)
    int n_sample, n_comp;
    const double* sample_times;
    double* samples;                // [n_traj][T][C]; slot 0 of a row is NaN when the trajectory never got there
    // work queue + counters
    unsigned long long* cursor;
    unsigned long long* counters;   // [CT_N]
    // record pool
    unsigned char* pool;            // pool_blocks * REC_BLOCK records of REC_WORDS 8-byte words
    unsigned long long pool_blocks;
    unsigned long long* blk_cursor; // blocks handed out (keeps counting past capacity)
    unsigned int* seg_table;        // [n_traj][table_w] first block of each segment
    int table_w;
    double* rate_pool;              // save_rate: total rate per pool record slot (null: not saved)
    unsigned long long save_c_mask, save_d_mask;   // ind_save_c / ind_save_d as bit masks (compaction only)
    int n_c_out, n_d_out;           // saved component counts = popcount of the masks
    int ev_pack;                    // event transport (xd_bytes = 1): the reaction index rides in the top byte of xd[0]'s
                                    // pool word (the library has checked |xd| < 2^23 for the whole run)
    // per-trajectory outputs
    unsigned long long* nrec;       // [n_traj + 1], records stored per trajectory
    long long* njumps;
    long long* nrejected;
    unsigned char* status;
};

// record layout in the pool: 8-byte words [t, xc[0..NC), xd packed 2 x int32], padded to an
// even word count so that every record is 16-byte aligned (vectorised stores)
template <class M>
struct RecLayout {
    static constexpr int XD_WORDS = (M::ND + 1) / 2;
    static constexpr int RAW = 1 + M::NC + XD_WORDS;
    static constexpr int WORDS = (RAW + 1) & ~1;
    static constexpr int BYTES = WORDS * 8;
};

// models whose rates depend on xd only may be run with the reference's ConstantRate wrapper
template <class M, class = void> struct ConstRateOk { static constexpr bool value = false; };
template <class M> struct ConstRateOk<M, decltype((void)M::CONST_RATE_OK)> { static constexpr bool value = M::CONST_RATE_OK; };

// models whose total rate is registered as a constant part + a variable part (reference CompositeRate)
template <class M, class = void> struct CompositeOk { static constexpr bool value = false; };
template <class M> struct CompositeOk<M, decltype((void)M::COMPOSITE_OK)> { static constexpr bool value = M::COMPOSITE_OK; };

// models whose CHV field is a sign vector times a scalar (TCP) use the single-chain RK attempt
template <class M, class = void> struct ChvSigned { static constexpr bool value = false; };
#ifndef PDMP_NO_SIGNED   // (experiment switch: the generic attempt for every model)
template <class M> struct ChvSigned<M, decltype((void)M::CHV_SIGNED_SCALAR)> { static constexpr bool value = M::CHV_SIGNED_SCALAR; };
#endif

// models that keep sub-expressions of the rates for the field at the post-jump state (`HAS_JUMP_CACHE`,
// `JumpCache<R>`, `rates_cached`, `chv_field_cached`): used when the jump leaves xc unchanged (no Delta)
template <class M, class = void> struct JumpCacheOk { static constexpr bool value = false; };
template <class M> struct JumpCacheOk<M, decltype((void)M::HAS_JUMP_CACHE)> {
#ifdef PDMP_NO_JUMP_CACHE
    static constexpr bool value = false;
#else
    static constexpr bool value = M::HAS_JUMP_CACHE && M::HAS_CHV_FIELD && !M::HAS_DELTA;
#endif
};

template <class M, class R, bool = JumpCacheOk<M>::value> struct JumpCacheOf { struct type {}; };
template <class M, class R> struct JumpCacheOf<M, R, true> { using type = typename M::template JumpCache<R>; };

// models whose F does not depend on (xc, t) between jumps (TCP: F = +-1): the plain-F flow is exact
template <class M, class = void> struct ConstFlow { static constexpr bool value = false; };
template <class M> struct ConstFlow<M, decltype((void)M::CONST_FLOW)> { static constexpr bool value = M::CONST_FLOW; };

// models that register the EXACT flow of F between jumps (`HAS_EXACT_FLOW`, `flow(x, xd, p, t, dt)`): sample times and
// the last bit are then one closed-form evaluation instead of an adaptive sub-integration (tcp2d: x' = 1, x' = -x)
template <class M, class = void> struct ExactFlow { static constexpr bool value = false; };
template <class M> struct ExactFlow<M, decltype((void)M::HAS_EXACT_FLOW)> { static constexpr bool value = M::HAS_EXACT_FLOW; };

// geometric segment sizes: segment k holds REC_BLOCK << (k >> 2) records (4 segments per doubling)
__host__ __device__ inline unsigned seg_records(int k) { return (unsigned)REC_BLOCK << (k >> 2); }

// R is the real type of the state and of the stage arithmetic: double, or float for the optional
// PDMP_PREC_FP32 kernels.  Time-like scalars (s, thresholds, step size, jump times) stay double
// in both; in FP32 CHV kernels the time slot y[NC] is refreshed after every accepted step from a
// double shadow `tacc` that accumulates the step increments, so that time does not drift.
template <class M, class Tab, int ALG, bool RECORDS, bool STATS, class R>
struct Traj {
    static constexpr int NC = M::NC, ND = M::ND, NR = M::NR;
    static constexpr int D = (ALG == PDMP_ALG_CHV) ? NC + 1 : NC;
    static constexpr bool SHADOW_T = (ALG == PDMP_ALG_CHV) && sizeof(R) == 4;
    using RL = RecLayout<M>;

    // ---- integrator state ----
    R y[D];             // CHV: (xc, t) in the time-change variable s ; Rejection: xc in time t
    R k1[D];            // FSAL
    double tacc;        // SHADOW_T only: the time slot in double
    R inv_rate;         // ConstantRate (src/rate.jl:17-40): 1 / total rate, cached from jump to jump;
                        // CompositeRate (src/rate.jl:42-69): the constant part of the total rate
    double r, tstop, h; // distance of the integration variable to the next threshold (s = tstop - r; an
                        // accepted step subtracts from it, a landing makes it exactly 0), the threshold, the
                        // proposed step.  The CHV field is autonomous in s, so s itself is never formed there.
    PIController pi;
    // ---- PDMP state ----
    int xd[ND];
    double tlast;       // reference `t` of the outer loop (time of the last threshold)
    R xs[NC]; double ts;  // anchor of the plain-F flow: state/time at the last executed jump
                        // (CHV: caract.xc ; Rejection: state at the last candidate)
    long long simnj, nfict;
    unsigned attempts;
    DrawStream rng;
    unsigned lid;       // local trajectory index
    // ---- writer ----
    unsigned nrec_stored, seg_left;
    int seg_k;
    unsigned long long wslot;  // next record slot in the pool
    bool overflow;
    int next_sample;    // index of the next sample time not yet recorded

    // ------------------------------------------------------------------------------
    PDMP_DEV double draw(const KParams& P, bool& exhausted) {
        return rng.next(P.rng_mode, P.replay_u + (size_t)lid * P.replay_stride, P.replay_stride,
                        P.traj_id_offset + lid, P.keys, exhausted);
    }

    // (chv::CHV)(xdot, x, caract, t): (F, 1) / Rtot      reference src/chvdiffeq.jl:6-16
    PDMP_DEV void field(const KParams& P, const R* x, R* dx, R t) const {
        if constexpr (ALG == PDMP_ALG_CHV) {
            if constexpr (M::HAS_CHV_FIELD) {
                M::chv_field(dx, x, xd, P.parms);
            } else {
                const R tau = x[NC];
                R inv;
                if constexpr (ConstRateOk<M>::value) inv = P.rate_const ? inv_rate : rcp_fast(M::rtot(x, xd, P.parms, tau));
                else if constexpr (CompositeOk<M>::value)   // out_cst .+ out_var, src/rate.jl:65-67 (inv_rate caches out_cst)
                    inv = rcp_fast(P.rate_composite ? inv_rate + M::rtot_var(x, xd, P.parms, tau) : M::rtot(x, xd, P.parms, tau));
                else inv = rcp_fast(M::rtot(x, xd, P.parms, tau));
                M::F(dx, x, xd, P.parms, tau);
#pragma unroll
                for (int i = 0; i < NC; ++i) dx[i] *= inv;
                dx[NC] = inv;
            }
        } else {
            M::F(dx, x, xd, P.parms, t);  // Rejection flow, reference src/rejectiondiffeq.jl:5-8
        }
    }

    // ConstantRate: cr.totalrate is recomputed at the first call after a jump (src/rate.jl:28-33)
    PDMP_DEV void refresh_rate(const KParams& P, double t) {
        if constexpr (ConstRateOk<M>::value && ALG == PDMP_ALG_CHV)
            if (P.rate_const) inv_rate = rcp_fast(M::rtot(y, xd, P.parms, (R)t));
        if constexpr (CompositeOk<M>::value && ALG == PDMP_ALG_CHV)
            if (P.rate_composite) inv_rate = M::rtot_cst(y, xd, P.parms, (R)t);   // the ConstantRate part only
    }

    // ---- record pool writer ----------------------------------------------------------
    PDMP_DEV static double no_rate() { return __longlong_as_double(0x7ff8000000000000ll); }   // NaN: no jump produced this point
    PDMP_DEV void write_record(const KParams& P, double t, const R* xc) { write_record(P, t, xc, no_rate(), 0); }
    // ev: the reaction (1-based) that changed xd since this trajectory's previous saved point, 0 if none
    PDMP_DEV void write_record(const KParams& P, double t, const R* xc, double rate_tot, int ev) {
        if constexpr (!RECORDS) return;
        if (seg_left == 0) {
            const unsigned nrecs = seg_records(seg_k);
            const unsigned nblk = nrecs / REC_BLOCK;
            const unsigned long long b = atomicAdd(P.blk_cursor, (unsigned long long)nblk);
            if (b + nblk > P.pool_blocks || seg_k >= P.table_w) overflow = true;
            if (!overflow) {
                P.seg_table[(size_t)lid * P.table_w + seg_k] = (unsigned)b;
                wslot = b * REC_BLOCK;
            }
            seg_left = nrecs;
            ++seg_k;
        }
        --seg_left;
        if (overflow) return;
        double w[RL::WORDS];
        w[0] = t;
#pragma unroll
        for (int i = 0; i < NC; ++i) w[1 + i] = (double)xc[i];
#pragma unroll
        for (int i = 0; i < RL::XD_WORDS; ++i) {
            int lo = xd[2 * i];
            if (i == 0 && P.ev_pack) lo = (lo & 0x00ffffff) | (ev << 24);
            const int hi = (2 * i + 1 < ND) ? xd[2 * i + 1] : 0;
            w[1 + NC + i] = __hiloint2double(hi, lo);
        }
#pragma unroll
        for (int i = RL::RAW; i < RL::WORDS; ++i) w[i] = 0.0;
        double2* dst = reinterpret_cast<double2*>(P.pool + wslot * RL::BYTES);
#pragma unroll
        for (int i = 0; i < RL::WORDS / 2; ++i) __stcs(dst + i, make_double2(w[2 * i], w[2 * i + 1]));
        if (P.rate_pool) __stcs(P.rate_pool + wslot, rate_tot);   // save_rate (reference rate_hist)
        ++wslot;
        ++nrec_stored;
    }

    // ---- the state at a sample time: one row of the sample buffer ------------------------------
    PDMP_DEV void accumulate(const KParams& P, int j, const R* xc) {
        if constexpr (!STATS) return;
        double* row = P.samples + ((size_t)lid * P.n_sample + j) * (NC + ND);
#pragma unroll
        for (int c = 0; c < NC; ++c) __stcs(row + c, (double)xc[c]);
#pragma unroll
        for (int c = 0; c < ND; ++c) __stcs(row + NC + c, (double)xd[c]);
    }

    // ---- plain-F flow of (x, t) to tau, either direction; adaptive, same pair, user tolerances ----
    // quirk = true: OrdinaryDiffEq's defaults of a fresh `solve(prob, ode)` — reltol 1e-3, abstol 1e-6 and the
    // Hairer-Norsett-Wanner initial step — as the reference's last bit does (src/chvdiffeq.jl:162-163)
    PDMP_DEV bool flow_F(const KParams& P, R (&x)[NC], double& t, double tau, unsigned& aux_attempts, int& fail, bool quirk = false) {
        if constexpr (M::ZERO_FLOW) { t = tau; return true; }
        if (t == tau) return true;
        if constexpr (ConstFlow<M>::value) {   // F is constant between jumps: the flow is x + F (tau - t), exactly
            R f[NC];
            M::F(f, x, xd, P.parms, (R)t);
#pragma unroll
            for (int i = 0; i < NC; ++i) x[i] = fma(f[i], (R)(tau - t), x[i]);
            t = tau;
            ++aux_attempts;
            return true;
        }
        if constexpr (ExactFlow<M>::value)
            if (!quirk) {   // (the reference-faithful last bit is by definition a loose numerical solve)
                M::flow(x, xd, P.parms, (R)t, (R)(tau - t));
                t = tau;
                ++aux_attempts;
                return true;
            }
        const double dir = tau > t ? 1.0 : -1.0;
        double hh = tau - t;
        const R rtol = quirk ? R(1e-3) : (R)P.reltol, atol = quirk ? R(1e-6) : (R)P.abstol;
        PIController c2; c2.init();
        R f1[NC];
        M::F(f1, x, xd, P.parms, (R)t);
        auto rhs = [&](const R* xx, R* dx, R tt) { M::F(dx, xx, xd, P.parms, tt); };
        if (quirk) {   // ode_determine_initdt, order 5, dtmax = the interval
            double d0 = 0.0, d1 = 0.0, sk[NC];
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                sk[i] = (double)atol + fabs((double)x[i]) * (double)rtol;
                const double a = (double)x[i] / sk[i], b = (double)f1[i] / sk[i];
                d0 = fma(a, a, d0); d1 = fma(b, b, d1);
            }
            d0 = sqrt(d0 * (1.0 / NC)); d1 = sqrt(d1 * (1.0 / NC));
            double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
            dt0 = fmin(dt0, fabs(hh));
            R u1[NC], f2[NC];
#pragma unroll
            for (int i = 0; i < NC; ++i) u1[i] = fma((R)dt0, f1[i], x[i]);
            M::F(f2, u1, xd, P.parms, (R)(t + dt0));
            double d2 = 0.0;
#pragma unroll
            for (int i = 0; i < NC; ++i) { const double a = ((double)f2[i] - (double)f1[i]) / sk[i]; d2 = fma(a, a, d2); }
            d2 = sqrt(d2 * (1.0 / NC)) / dt0;
            const double dm = fmax(d1, d2);
            const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(dm)) / 5.0);
            hh = fmin(fmin(100.0 * dt0, dt1), fabs(hh));
        }
        const unsigned budget = (unsigned)min((long long)0x7fffffff, P.max_attempts);
        while (t != tau) {
            bool clipped = false;
            double step = hh;
            if (dir * step >= dir * (tau - t)) { step = tau - t; clipped = true; }
            R xn[NC], f7[NC], inc;
            const R e2 = rk_attempt<Tab, NC, R>(rhs, x, f1, (R)t, (R)step, xn, f7, rtol, atol, inc);
            if (++aux_attempts > budget) { fail = PDMP_ST_MAX_ATTEMPTS; return false; }
            const bool accept = e2 <= (R)NC;
            double fac;
            if (!(e2 == e2) || isinf(e2)) {
                bool ok = true;
#pragma unroll
                for (int i = 0; i < NC; ++i) ok = ok && isfinite(x[i]) && isfinite(f1[i]);
                if (!ok || !(t + 0.2 * step != t)) { fail = PDMP_ST_NAN; return false; }
                fac = 0.2;
            } else {
                fac = c2.step((float)(e2 * (R)(1.0 / NC)), accept);
            }
            if (accept) {
#pragma unroll
                for (int i = 0; i < NC; ++i) { x[i] = xn[i]; f1[i] = f7[i]; }
                double tn = t + step;
                if (clipped || fabs(tn - tau) < 100.0 * 2.220446049250313e-16 * fmax(fabs(tn), fabs(tau))) tn = tau;
                t = tn;
            }
            hh = step * fac;
        }
        return true;
    }
    // the "last bit": plain-F flow of the anchor (state at the last executed jump) to tf,
    // reference src/chvdiffeq.jl:160-168 (a fresh ODE solve there)
    PDMP_DEV bool flow_anchor_to(const KParams& P, double tau, unsigned& aux_attempts, int& fail) {
        if (!(tau > ts)) return true;
        return flow_F(P, xs, ts, tau, aux_attempts, fail, P.ref_quirks != 0);
    }

    // ---- ensemble statistics at the fixed sample times -----------------------------------
    // Lanes never stop for a sample time.  At the next threshold (before the jump) every sample time passed since
    // the previous one is reached by flowing plain F BACKWARDS from the pre-jump state — xd has not changed in
    // between — latest first, so that every leg is short: exactly for models that register their flow
    // (ExactFlow / ConstFlow), by an adaptive sub-integration otherwise.
    // time the trajectory has reached (CHV: the time slot; Rejection: s is time)
    PDMP_DEV double svar() const { return tstop - r; }   // the integration variable (Rejection: time)
    PDMP_DEV double threshold_time() const {
        if constexpr (SHADOW_T) return tacc;
        else if constexpr (ALG == PDMP_ALG_CHV) return (double)y[D - 1];
        else return svar();
    }
    PDMP_DEV bool sample_due(const KParams& P) const {
        if constexpr (!STATS) return false;
        return next_sample < P.n_sample && __ldg(P.sample_times + next_sample) <= threshold_time();   // (sample times are <= tf by contract)
    }
    PDMP_DEV int pre_sample(const KParams& P, unsigned& aux) {
        int fail = PDMP_ST_OK;
        if constexpr (STATS) {
            const double tcur = threshold_time(), tlim = fmin(tcur, P.tf);
            int last = next_sample;
            while (last < P.n_sample && __ldg(P.sample_times + last) <= tlim) ++last;
            R xb[NC];
            double tb = tcur;
#pragma unroll
            for (int i = 0; i < NC; ++i) xb[i] = y[i];
            for (int j = last - 1; j >= next_sample; --j) {      // latest first: always a short backward leg
                if (!flow_F(P, xb, tb, __ldg(P.sample_times + j), aux, fail)) return fail;
                accumulate(P, j, xb);
            }
            next_sample = last;
        }
        return fail;
    }

    // ---- init!(problem) + integrator construction --------------------------------------
    // reference src/problem.jl:150-161, src/chvdiffeq.jl:97-136, src/rejectiondiffeq.jl:91-116
    PDMP_DEV void init(const KParams& P, unsigned idx, bool& exhausted) {
        lid = idx;
        rng.init();
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            y[i] = (R)P.xc0[(P.xc0_per_traj ? (size_t)idx * NC : 0) + i];
            xs[i] = y[i];
        }
#pragma unroll
        for (int i = 0; i < ND; ++i) xd[i] = (int)P.xd0[(P.xd0_per_traj ? (size_t)idx * ND : 0) + i];
        const double E0 = neg_log_unit(draw(P, exhausted));  // tstop_extended = -log(rand())
        double s0;
        if constexpr (ALG == PDMP_ALG_CHV) {
            y[NC] = (R)P.t0;  // X_extended[end] = ti
            s0 = 0.0;
            tstop = E0;
        } else {
            s0 = P.t0;
            tstop = E0 / (double)M::bound(y, xd, P.parms, (R)P.t0) + P.t0;  // src/rejectiondiffeq.jl:105-107
        }
        r = tstop - s0;
        tacc = P.t0;
        tlast = P.t0; ts = P.t0;
        simnj = 1; nfict = 0;
        attempts = 0;
        nrec_stored = 0; seg_left = 0; seg_k = 0; wslot = 0; overflow = false;
        next_sample = 0;
        pi.init();
        write_record(P, P.t0, xs);  // the initial point is always kept (resize!(..., 1))

        // Hairer-Norsett-Wanner initial step (OrdinaryDiffEq ode_determine_initdt), order 5
        refresh_rate(P, P.t0);
        field(P, y, k1, (R)s0);
        double d0 = 0.0, d1 = 0.0;
        double sk[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            sk[i] = P.abstol + fabs((double)y[i]) * P.reltol;
            const double a = (double)y[i] / sk[i], b = (double)k1[i] / sk[i];
            d0 = fma(a, a, d0); d1 = fma(b, b, d1);
        }
        d0 = sqrt(d0 * (1.0 / D)); d1 = sqrt(d1 * (1.0 / D));
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
        dt0 = fmin(dt0, 1e9);
        R u1[D], f1[D];
#pragma unroll
        for (int i = 0; i < D; ++i) u1[i] = fma((R)dt0, k1[i], y[i]);
        field(P, u1, f1, (R)(s0 + dt0));
        double d2 = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) { const double a = ((double)f1[i] - (double)k1[i]) / sk[i]; d2 = fma(a, a, d2); }
        d2 = sqrt(d2 * (1.0 / D)) / dt0;
        const double dm = fmax(d1, d2);
        // 10^(-(2 + log10 dm)/5) = (100 dm)^(-1/5), evaluated in FP32 (it only seeds the controller)
        const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : (double)exp2f(-0.2f * __log2f((float)(100.0 * dm)));
        h = fmin(fmin(100.0 * dt0, dt1), 1e9);
    }

    PDMP_DEV void finish(const KParams& P, int st) {
        if (overflow && (st == PDMP_ST_OK || st == PDMP_ST_HIT_NJUMPS)) st = PDMP_ST_OUTPUT_OVERFLOW;
        if constexpr (STATS)   // sample times this trajectory never reached (cap, failure): not counted
            for (int j = next_sample; j < P.n_sample; ++j)
                __stcs(P.samples + ((size_t)lid * P.n_sample + j) * (NC + ND), __longlong_as_double(0x7ff8000000000000ll));
        P.nrec[lid] = nrec_stored;
        P.njumps[lid] = simnj;
        P.nrejected[lid] = nfict;
        P.status[lid] = (unsigned char)st;
    }

    // ---- loop epilogue shared by both algorithms: end-of-trajectory tests ----------------
    // returns true when the trajectory is finished
    PDMP_DEV bool epilogue(const KParams& P, bool rng_exhausted, unsigned& aux, int& fail) {
        if (rng_exhausted) { finish(P, PDMP_ST_REPLAY_EXHAUSTED); return true; }
        if (!(tlast < P.tf)) {  // while (t < tf) ...
            if (tlast > P.tf) {  // last bit
                if (!flow_anchor_to(P, P.tf, aux, fail)) { finish(P, fail); return true; }
                write_record(P, P.tf, xs);
            }
            finish(P, PDMP_ST_OK);
            return true;
        }
        if (!(simnj < P.n_jumps)) { finish(P, PDMP_ST_HIT_NJUMPS); return true; }
        return false;
    }

    // ---- chvjump, reference src/chvdiffeq.jl:24-74, and the tail of the while body :145-157 ----
    // (the sample times at or before this threshold were handled by pre_sample)
    PDMP_DEV bool chv_threshold(const KParams& P, unsigned& aux, unsigned& jumps) {
        int fail = PDMP_ST_OK;
        bool ex = false;
        const double tj = threshold_time();
        R rate[NR];
        [[maybe_unused]] typename JumpCacheOf<M, R>::type jc;
        if constexpr (JumpCacheOk<M>::value) M::rates_cached(rate, y, xd, P.parms, (R)tj, jc);
        else M::rates(rate, y, xd, P.parms, (R)tj);  // :51, rates at the PRE-jump state
        R tot = rate[0];
#pragma unroll
        for (int i = 1; i < NR; ++i) tot += rate[i];  // sum(rate), left to right
        if (P.save_pre && tj <= P.tf) write_record(P, tj, y, (double)tot, 0);  // :41-48
        double u_exp;
        int ev_done = 0;
        if (tj < P.tf) {  // :52
            // a positive finite total is required where pfsample uses it; at the overshoot threshold the
            // reference computes the rates and ignores them
            if (!(tot > R(0) && tot < (sizeof(R) == 4 ? (R)3.4028234e38f : (R)1.7976931348623157e308))) {
                finish(P, PDMP_ST_NONPOSITIVE_RATE); return true;
            }
            // rand() of pfsample then rand() of :71 = the cached half of the previous Philox block and the
            // first half of a new one: ONE Philox call, no parity branch (DrawStream::next2_odd)
            double u_ev;
            rng.next2_odd(P.rng_mode, P.replay_u + (size_t)lid * P.replay_stride, P.replay_stride, P.traj_id_offset + lid, P.keys,
                          ex, u_ev, u_exp);
            // pfsample, reference src/utils.jl:46-57 (strict <, last index absorbs round-off)
            const R thr = (R)u_ev * tot;
            int ev = 1;
            R cw = rate[0];
#pragma unroll
            for (int i = 1; i < NR; ++i)
                if (ev == i && cw < thr) { ev = i + 1; cw += rate[i]; }
            // affect!, reference src/jumps.jl:28-36
#pragma unroll
            for (int i = 0; i < ND; ++i) xd[i] += P.nu[(ev - 1) * ND + i];
            if constexpr (M::HAS_DELTA) M::delta(y, xd, P.parms, tj, ev);
            ev_done = ev;
            refresh_rate(P, tj);
            // u_modified!(integrator, true): FSAL re-evaluated (s = tstop here)
            if constexpr (JumpCacheOk<M>::value) M::chv_field_cached(k1, y, xd, P.parms, jc);
            else field(P, y, k1, (R)tstop);
#pragma unroll
            for (int i = 0; i < NC; ++i) xs[i] = y[i];  // caract.xc .= u
            ts = tj;
            ++jumps;
        } else {
            u_exp = draw(P, ex);          // the overshoot threshold draws its exponential only
        }
        ++simnj;                          // :70
        {
            const double tnew = tstop + neg_log_unit(u_exp);   // :71, always drawn
            r = tnew - tstop;                         // s sits exactly on the old threshold
            tstop = tnew;
        }
        if (!(tlast < tj)) { finish(P, PDMP_ST_NO_PROGRESS); return true; }  // @assert :145
        tlast = tj;                       // :146
        if (P.save_post && tj <= P.tf) write_record(P, tj, xs, tj < P.tf ? (double)tot : no_rate(), ev_done);  // :149-156
        return epilogue(P, ex, aux, fail);
    }

    // ---- rejectionjump, reference src/rejectiondiffeq.jl:14-74, and :121-139 ----
    PDMP_DEV bool rej_candidate(const KParams& P, unsigned& aux, unsigned& jumps) {
        int fail = PDMP_ST_OK;
        bool ex = false;
        const double tc = tstop;    // the candidate time: the lane landed, s == tstop
        R rate[NR];
        M::rates(rate, y, xd, P.parms, (R)tc);
        R tot = rate[0];
#pragma unroll
        for (int i = 1; i < NR; ++i) tot += rate[i];
        const R bnd = M::bound(y, xd, P.parms, (R)tc);
        if (!(tot < bnd)) { finish(P, PDMP_ST_BOUND_VIOLATED); return true; }  // @assert :33
        double u_acc, u_exp;   // rand() of :35 then rand() of :36, one converged Philox call for both
        rng.next2(P.rng_mode, P.replay_u + (size_t)lid * P.replay_stride, P.replay_stride, P.traj_id_offset + lid, P.keys,
                  ex, u_acc, u_exp);
        const bool reject = u_acc < 1.0 - (double)tot / (double)bnd;         // :35
        const double dtn = neg_log_unit(u_exp) / (double)bnd;                        // :36
        const bool fire = tc < P.tf && !reject;
        int ev_done = 0;
        if (fire) {                                          // :41
            const R thr = (R)draw(P, ex) * tot;              // :55
            int ev = 1;
            R cw = rate[0];
#pragma unroll
            for (int i = 1; i < NR; ++i)
                if (ev == i && cw < thr) { ev = i + 1; cw += rate[i]; }
#pragma unroll
            for (int i = 0; i < ND; ++i) xd[i] += P.nu[(ev - 1) * ND + i];
            if constexpr (M::HAS_DELTA) M::delta(y, xd, P.parms, tc, ev);
            ev_done = ev;
            field(P, y, k1, (R)tc);
            ++simnj;                                         // :64
            ++jumps;
        } else {
            ++nfict;                                         // :66
        }
        if (tc < P.tf) {
            // anchor of the last bit: the state at this candidate — or, the way the reference does it
            // (ref_quirks), caract.xc of the last ACCEPTED jump paired with the time of the last candidate
            if (fire || !P.ref_quirks) {
#pragma unroll
                for (int i = 0; i < NC; ++i) xs[i] = y[i];
            }
            ts = tc;
        }
        tstop = tc + dtn;                                    // :71
        r = tstop - tc;
        if (tlast >= tc) { finish(P, PDMP_ST_NO_PROGRESS); return true; }  // :121-124
        tlast = tc;
        if (P.save_post && tc <= P.tf && !reject) write_record(P, tc, y, (double)tot, ev_done);  // :128-138
        return epilogue(P, ex, aux, fail);
    }
};

// resident CTAs per SM the kernel is compiled for: small models run 6 x 128 threads (<= 80
// registers), those with a hand-simplified CHV field (TCP) 7 x 128 (<= 72 registers, no spills; measured on
// TCP with the final round-2 kernel, 4 M trajectories: 6 CTAs 34.4 ms, 7 CTAs 34.2 ms, 8 CTAs / 64 registers
// with 174 spilled bytes 35.4 ms, 9 CTAs 35.0 ms; 10^7 trajectories: 76.4 / 76.3 / 77.6 ms — earlier in the
// round, with a longer loop, 8 had been ahead), medium ones 6 (<= 80; tcp2d 4 / 5 / 6 CTAs: 16.5 / 14.95 / 14.63 ms), the larger ones 4 (<= 128
// registers; Morris-Lecar loses 10 % to spills at 96)
#ifndef PDMP_MINB_SMALL
#define PDMP_MINB_SMALL 7
#endif
#ifndef PDMP_MINB_MEDIUM
#define PDMP_MINB_MEDIUM 6
#endif
#ifndef PDMP_MINB_NOFLOW
#define PDMP_MINB_NOFLOW 8
#endif
#ifndef PDMP_MINB_LARGE
#define PDMP_MINB_LARGE 4
#endif
template <class M, int ALG, class R>
struct MinBlocks {
    static constexpr int D = (ALG == PDMP_ALG_CHV) ? M::NC + 1 : M::NC;
    static constexpr bool small = (D <= 2 && M::NR <= 2 && M::ND <= 2);
    static constexpr bool medium = (D <= 3 && M::NR <= 2 && M::ND <= 2);   // tcp2d: 96 registers, measured +8 %
    // Rejection with F = F_dummy: no RK state at all, the kernel only runs the candidate section and is latency-bound
    // (Philox -> log -> two divisions): as many resident warps as the register file gives
    static constexpr bool no_flow = (ALG == PDMP_ALG_REJECTION) && M::ZERO_FLOW;
    static constexpr int value = no_flow ? PDMP_MINB_NOFLOW
                                 : sizeof(R) == 4 ? (small ? 7 : 5) : (small ? (M::HAS_CHV_FIELD ? PDMP_MINB_SMALL : 6) : medium ? PDMP_MINB_MEDIUM : PDMP_MINB_LARGE);
};

// lane states of the phase scheduler
// (LANE_FAIL + status: the RK phase found the trajectory dead — non-finite state, attempt budget — and leaves
//  the once-per-trajectory exit to the threshold phase, so that none of its code sits in the RK loop)
enum { LANE_DEAD = 0, LANE_RUN = 1, LANE_EXHAUSTED = 2, LANE_WAIT = 3, LANE_FAIL = 4 };

// default scheduler thresholds; a model may define `static constexpr int SCHED_JUMP, SCHED_REFILL`
template <class M, class = void> struct SchedDefaults { static constexpr int jump = 12, refill = 2; };
template <class M> struct SchedDefaults<M, decltype((void)M::SCHED_JUMP, (void)M::SCHED_REFILL)> {
    static constexpr int jump = M::SCHED_JUMP, refill = M::SCHED_REFILL;
};

template <class M, class Tab, int ALG, bool RECORDS, bool STATS, class R>
__global__ void __launch_bounds__(BLOCK, MinBlocks<M, ALG, R>::value) ensemble_kernel(const __grid_constant__ KParams P) {
    using T = Traj<M, Tab, ALG, RECORDS, STATS, R>;
    constexpr int D = T::D;
    constexpr bool NO_FLOW = (ALG == PDMP_ALG_REJECTION && M::ZERO_FLOW);

    T tr;
    int st = LANE_DEAD;
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long c_attempts = 0;
    unsigned c_rejects = 0, c_aux = 0, c_jumps = 0, c_cands = 0;
    const unsigned budget = (unsigned)min((long long)0x7fffffff, P.max_attempts);

    // Phase scheduler.  Every iteration the warp does ONE of three things, chosen warp-uniformly:
    //   RK     one adaptive step attempt for every RUN lane (the arithmetic-heavy, converged part)
    //   JUMP   the threshold section for every WAIT lane — deferred until enough lanes wait, so
    //          that its ~300 instructions are issued for many lanes at once instead of ~5
    //   REFILL trajectories for DEAD lanes from the global cursor (warp-aggregated atomic)
    // Greedy rule behind the JUMP threshold: with w waiting and r running lanes, landing probability
    // p per lane and RK iteration, n iterations since the last JUMP and phase costs J (JUMP) and R (one
    // RK iteration), one more RK iteration pays iff p r (n + J/R) > w.  For TCP (p = 0.18, J/R = 1.5)
    // that gives w ~ 14-16, for Morris-Lecar (p = 0.5, an RK iteration of 18 exp) w ~ 6: hence the
    // per-model defaults SchedDefaults<M> (measured sweeps: profiles/r1_sched_sweep.txt).
    for (;;) {
        const unsigned m_run = __ballot_sync(0xffffffffu, st == LANE_RUN);
        const unsigned m_wait = __ballot_sync(0xffffffffu, st >= LANE_WAIT);
        const unsigned m_dead = __ballot_sync(0xffffffffu, st == LANE_DEAD);
        if (!(m_run | m_wait | m_dead)) break;
        const int n_run = __popc(m_run), n_wait = __popc(m_wait), n_dead = __popc(m_dead);

        if (n_dead && (n_dead >= P.refill_thresh || n_run == 0)) {
            // ---- REFILL ----
            unsigned long long base = 0;
            const int leader = __ffs(m_dead) - 1;
            if ((int)lane == leader) base = atomicAdd(P.cursor, (unsigned long long)n_dead);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (st == LANE_DEAD) {
                const unsigned long long idx = base + __popc(m_dead & ((1u << lane) - 1u));
                if (idx < P.n_traj) {
                    bool ex = false;
                    tr.init(P, (unsigned)idx, ex);
                    // while (t < tf) && (njumps < n_jumps): may be false before the first step
                    if (!(tr.simnj < P.n_jumps)) tr.finish(P, PDMP_ST_HIT_NJUMPS);
                    else if (ex) tr.finish(P, PDMP_ST_REPLAY_EXHAUSTED);
                    else if (NO_FLOW) { tr.r = 0.0; st = LANE_WAIT; }
                    else st = LANE_RUN;
                } else {
                    st = LANE_EXHAUSTED;
                }
            }
            continue;
        }

        const int n_live = n_run + n_wait;
        // Enough lanes wait (or none can run): straight to the threshold section.  Otherwise an RK phase first; it
        // ends when enough lanes have landed, so the threshold section ALWAYS follows it — one scheduling decision
        // per (RK, JUMP) pair instead of one per phase.
        const int jcap = (n_live * P.jump_cap_q + 3) >> 2;
        const bool jump_now = n_wait && (n_run == 0 || n_wait >= min(P.jump_thresh, jcap));
        if (!jump_now) {
        // ---- RK: adaptive attempts, clipped at the threshold (tstops), until enough lanes wait ----
        // Lanes only LEAVE the running set in this phase, so the scheduler's decision reduces to one
        // ballot per attempt; the full three-way test is redone when the threshold is reached.
        // (compiled out when there is no flow: its registers would cap the occupancy of a kernel that
        //  only ever runs the candidate section)
        if constexpr (!NO_FLOW) {
        const int thr = P.jump_thresh > 0 ? min(P.jump_thresh, jcap) : 1;
        // the phase ends when enough lanes wait for the threshold section (or none runs): one ballot per attempt
        const int run_floor = max(n_run - max(thr - n_wait, 1), 0);
        // OrdinaryDiffEq snaps to a tstop that is within 100 eps of the step's end
        const double snaptol = 2.220446049250313e-14 * fabs(tr.tstop);
        const bool pol1 = P.tstop_policy == 1;
        for (;;) {
        if (st == LANE_RUN) {
            const double hprop = tr.h;
            const bool clipped = hprop >= tr.r;              // tstops: the step ends on the threshold
            const double step = clipped ? tr.r : hprop;
            R yn[D], k7[D], inc, e2;
            if constexpr (ALG == PDMP_ALG_CHV && ChvSigned<M>::value) {
                R g7;
                const unsigned neg = M::chv_negmask(tr.xd, P.parms);
                e2 = rk_attempt_signed<Tab, D, R>([&](const R* x) { return M::chv_scalar(x, tr.xd, P.parms); }, neg, tr.y,
                                                  tr.k1[D - 1], R(0), (R)step, yn, g7, (R)P.reltol, (R)P.abstol, inc);
#pragma unroll
                for (int i = 0; i < D; ++i) k7[i] = apply_sign(g7, neg, i);
            } else {
                auto rhs = [&](const R* x, R* dx, R t) { tr.field(P, x, dx, t); };
                e2 = rk_attempt<Tab, D, R>(rhs, tr.y, tr.k1, (R)tr.svar(), (R)step, yn, k7, (R)P.reltol, (R)P.abstol, inc);
            }
            ++tr.attempts;
            if (e2 <= (R)D) {                                // accepted (false for a NaN estimate)
#pragma unroll
                for (int i = 0; i < D; ++i) { tr.y[i] = yn[i]; tr.k1[i] = k7[i]; }
                if constexpr (T::SHADOW_T) { tr.tacc += (double)inc; tr.y[D - 1] = (R)tr.tacc; }
                double hnext = step * (double)tr.pi.template accept<D>((float)e2);
                const double r2 = tr.r - step;
                // threshold reached: exactly (clipped step) or within 100 eps (the floating-point snap)
                const bool land = clipped || r2 < snaptol;
                tr.r = land ? 0.0 : r2;
                if (land) st = LANE_WAIT;
                if (pol1 && clipped && hnext < hprop) hnext = hprop;
                tr.h = hnext;
            } else {
                ++c_rejects;
                if (e2 < (sizeof(R) == 4 ? (R)3.4028234e38f : (R)1.7976931348623157e308)) {
                    tr.h = step * (double)tr.pi.template reject<D>((float)e2);
                } else {
                    // trial step left the domain of F/R (non-finite estimate): reject with the maximal shrink,
                    // what OrdinaryDiffEq does with an infinite estimate; fatal only if the accepted state
                    // itself is non-finite or the step no longer advances s
                    bool ok = true;
#pragma unroll
                    for (int i = 0; i < D; ++i) ok = ok && isfinite(tr.y[i]) && isfinite(tr.k1[ChvSigned<M>::value ? D - 1 : i]);
                    tr.h = 0.2 * step;
                    const double sv = tr.svar();
                    if (!ok || !(sv + tr.h > sv)) st = LANE_FAIL + PDMP_ST_NAN;
                }
            }
            // OrdinaryDiffEq's maxiters: the budget is spent and the threshold is still ahead
            if (st == LANE_RUN && tr.attempts >= budget) st = LANE_FAIL + PDMP_ST_MAX_ATTEMPTS;
        }
        if (__popc(__ballot_sync(0xffffffffu, st == LANE_RUN)) <= run_floor) break;
        }
        }
        }
        {
            // ---- JUMP: the threshold section for every waiting lane (and the exit of lanes that failed in RK) ----
            // sampling pass first: the few lanes whose threshold lies beyond a fixed sample time flow
            // their anchor there; a block of its own, so that the warp is converged again for the jump pass
            int fc = 0;
            if constexpr (STATS) {
                const bool due = st == LANE_WAIT && tr.sample_due(P);
                if (__any_sync(0xffffffffu, due)) {
                    if (due) fc = tr.pre_sample(P, c_aux);
                }
            }
            if (st >= LANE_WAIT) {
                bool done;
                if (st > LANE_WAIT) fc = st - LANE_FAIL;     // failed in the RK phase
                if (fc) { tr.finish(P, fc); done = true; }
                else {
                    ++c_cands;
                    if constexpr (ALG == PDMP_ALG_CHV) done = tr.chv_threshold(P, c_aux, c_jumps);
                    else done = tr.rej_candidate(P, c_aux, c_jumps);
                }
                if (done) { st = LANE_DEAD; c_attempts += tr.attempts; }
                else if (NO_FLOW) { tr.r = 0.0; }   // F = F_dummy: next candidate directly
                else st = LANE_RUN;
            }
        }
    }

    // ---- flush counters (warp reduce, one atomic per warp) and CTA statistics ----
    unsigned long long v[CT_N] = {c_attempts, c_rejects, c_aux, c_jumps, c_cands};
#pragma unroll
    for (int k = 0; k < CT_N; ++k) {
        unsigned long long x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && x) atomicAdd(P.counters + k, x);
    }
}


// -----------------------------------------------------------------------------------------
// Host-visible launcher table entry
// -----------------------------------------------------------------------------------------
struct ModelEntry {
    pdmp_model_info info;
    int rec_words;  // RecLayout::WORDS
    int sched_jump, sched_refill;  // SchedDefaults<M>
    int const_rate_ok;             // ConstRateOk<M>
    // launch(alg, ode, precision, records, stats, P, grid, smem, stream): returns cudaError_t
    cudaError_t (*launch)(int alg, int ode, int prec, bool records, bool stats, const KParams& P, int grid, size_t smem,
                          cudaStream_t stream);
    // occupancy: resident CTAs per SM for that variant
    cudaError_t (*occupancy)(int alg, int ode, int prec, bool records, bool stats, size_t smem, int* blocks_per_sm);
    // compaction pool -> CSR
    cudaError_t (*compact)(const KParams& P, const unsigned long long* offsets, double* time, double* xc,
                           void* xd, double* rate, int xd_bytes, bool co_resident, cudaStream_t stream);
    int composite_ok;              // CompositeOk<M>
};

// pool -> CSR: one warp per trajectory walks its segments; reads are 16-byte vector loads of
// whole records, writes are contiguous runs per output array (time / xc / xd widened to int64)
// Launch shape: 128 threads and at most 32 registers per thread, so that one compaction CTA
// fits in the register space the resident ensemble CTAs leave free on every SM (e.g. TCP:
// 6 x 128 x 80 = 61440 of 65536 registers) and the two kernels truly run concurrently.
#ifndef PDMP_COMPACT_CTAS
#define PDMP_COMPACT_CTAS 8
#endif
#ifndef PDMP_COMPACT_UNROLL
#define PDMP_COMPACT_UNROLL 2
#endif
// one saved point: pool record -> the three CSR arrays
template <class M, class XT>
PDMP_DEV void csr_store(const double (&w)[RecLayout<M>::WORDS], unsigned long long o, double* __restrict__ time,
                        double* __restrict__ xc, XT* __restrict__ xd) {
    constexpr int NC = M::NC, ND = M::ND;
    __stcs(time + o, w[0]);
#pragma unroll
    for (int c = 0; c < NC; ++c) __stcs(xc + o * NC + c, w[1 + c]);
    if constexpr (sizeof(XT) == 1) {
        // event transport: ONE byte per saved point, the reaction index (the bindings rebuild xd by a prefix sum over nu)
        __stcs(reinterpret_cast<unsigned char*>(xd) + o, (unsigned char)((unsigned)__double2loint(w[1 + NC]) >> 24));
    } else if constexpr (sizeof(XT) == 2 && ND % 2 == 0) {
        // int16 transport: a pair of components is one 32-bit store
        unsigned* dst = reinterpret_cast<unsigned*>(xd + o * ND);
#pragma unroll
        for (int c = 0; c < ND; c += 2) {
            const double pw = w[1 + NC + c / 2];
            __stcs(dst + c / 2, ((unsigned)__double2loint(pw) & 0xffffu) | ((unsigned)__double2hiint(pw) << 16));
        }
    } else if constexpr (sizeof(XT) == 4 && ND % 2 == 0) {
        // int32 transport: the packed pool word is already the CSR layout
        double* dst = reinterpret_cast<double*>(xd + o * ND);
#pragma unroll
        for (int c = 0; c < ND; c += 2) __stcs(dst + c / 2, w[1 + NC + c / 2]);
    } else {
#pragma unroll
        for (int c = 0; c < ND; ++c) {
            const double pw = w[1 + NC + c / 2];
            const int v = (c & 1) ? __double2hiint(pw) : __double2loint(pw);
            xd[o * ND + c] = (XT)v;
        }
    }
}

// the same with ind_save_c / ind_save_d masks (reference src/chv.jl:54-62): only the selected components, in
// increasing index order, nco / ndo columns per saved point
template <class M, class XT>
PDMP_DEV void csr_store_masked(const double (&w)[RecLayout<M>::WORDS], unsigned long long o, double* __restrict__ time,
                               double* __restrict__ xc, XT* __restrict__ xd, unsigned long long cmask, unsigned long long dmask,
                               int nco, int ndo) {
    constexpr int NC = M::NC, ND = M::ND;
    __stcs(time + o, w[0]);
    int k = 0;
#pragma unroll
    for (int c = 0; c < NC; ++c)
        if ((cmask >> c) & 1ull) { xc[o * nco + k] = w[1 + c]; ++k; }
    if constexpr (sizeof(XT) == 1) {   // event transport: the reaction index, whatever the masks
        __stcs(reinterpret_cast<unsigned char*>(xd) + o, (unsigned char)((unsigned)__double2loint(w[1 + NC]) >> 24));
        return;
    }
    k = 0;
#pragma unroll
    for (int c = 0; c < ND; ++c)
        if ((dmask >> c) & 1ull) {
            const double pw = w[1 + NC + c / 2];
            xd[o * ndo + k] = (XT)((c & 1) ? __double2hiint(pw) : __double2loint(pw));
            ++k;
        }
}

// CO = true: the variant that shares the SMs with a running ensemble kernel (host-mode pipeline):
// 128 threads, <= 32 registers, one group of 32 records in flight per warp.  CO = false: the
// stand-alone variant of resident runs: two groups of 32 records in flight per warp (the kernel is
// bound by dependent-load latency, not by instructions), 8 CTAs per SM.  Measured pool -> CSR times
// (TCP 4 M / SIR 2 M / Morris-Lecar 1 M trajectories): co-resident shape 7.0 / 9.5 / 16.0 ms; unroll
// 2 x 8 CTAs 6.7 / 7.0 / 12.8; unroll 4 x 4 CTAs 6.7 / 6.2 / 10.0 — but the latter loses at the
// bench size (10^7 TCP trajectories, mostly short: 18.5 ms against 16.3 for 2 x 8 and 16.6 for the
// co-resident shape), where many warps in flight matter more than open DRAM pages.
template <class M, class XT, bool CO>
__global__ void __launch_bounds__(128, CO ? 16 : PDMP_COMPACT_CTAS) compact_kernel(const KParams P, const unsigned long long* __restrict__ offsets,
                                                               double* __restrict__ time, double* __restrict__ xc,
                                                               XT* __restrict__ xd, double* __restrict__ rate) {
    using RL = RecLayout<M>;
    constexpr int UN = CO ? 1 : PDMP_COMPACT_UNROLL;
    const bool masked = P.n_c_out != M::NC || P.n_d_out != M::ND;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long warp0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long i = warp0; i < P.n_traj; i += nwarps) {
        const unsigned n = (unsigned)P.nrec[i];
        const unsigned long long off = offsets[i];
        const unsigned* tab = P.seg_table + i * P.table_w;
        // record r lives in segment k = 4m + q / (16 << m) at position q % (16 << m), where
        // m = floor(log2(r / 64 + 1)) is the size level and q = r - 64 (2^m - 1): every warp
        // iteration moves 32 consecutive records whatever the segment boundaries
        for (unsigned r0 = lane; r0 < n; r0 += 32 * UN) {
            double w[UN][RL::WORDS];
            double rt[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const unsigned r = r0 + 32 * u;
                if (r < n) {
                    const unsigned m = 31u - __clz((r >> 6) + 1u);
                    const unsigned q = r - (64u << m) + 64u;
                    const unsigned k = 4u * m + (q >> (4u + m));
                    const unsigned pos = q & ((16u << m) - 1u);
                    const unsigned long long slot = (unsigned long long)__ldg(tab + k) * REC_BLOCK + pos;
                    if (rate) rt[u] = __ldcs(P.rate_pool + slot);
                    const double2* src = reinterpret_cast<const double2*>(P.pool + slot * RL::BYTES);
#pragma unroll
                    for (int j = 0; j < RL::WORDS / 2; ++j) {
                        const double2 v = __ldcs(src + j);
                        w[u][2 * j] = v.x; w[u][2 * j + 1] = v.y;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const unsigned r = r0 + 32 * u;
                if (r < n) {
                    if (masked) csr_store_masked<M, XT>(w[u], off + r, time, xc, xd, P.save_c_mask, P.save_d_mask, P.n_c_out, P.n_d_out);
                    else csr_store<M, XT>(w[u], off + r, time, xc, xd);
                    if (rate) __stcs(rate + off + r, rt[u]);
                }
            }
        }
    }
}

// calls f(kernel function pointer) for the requested variant
template <class M, class Tab, int ALG, class R, class Fn>
cudaError_t with_variant(bool records, bool stats, Fn&& f) {
    if (records && stats) return f(ensemble_kernel<M, Tab, ALG, true, true, R>);
    if (records) return f(ensemble_kernel<M, Tab, ALG, true, false, R>);
    if (stats) return f(ensemble_kernel<M, Tab, ALG, false, true, R>);
    return f(ensemble_kernel<M, Tab, ALG, false, false, R>);
}
template <class M, int ALG, class Fn>
cudaError_t with_alg_variant(int ode, int prec, bool records, bool stats, Fn&& f) {
    if (prec == PDMP_PREC_FP32) {
        return ode == PDMP_ODE_TSIT5 ? with_variant<M, TabTsit5, ALG, float>(records, stats, f)
                                     : with_variant<M, TabDP5, ALG, float>(records, stats, f);
    }
    return ode == PDMP_ODE_TSIT5 ? with_variant<M, TabTsit5, ALG, double>(records, stats, f)
                                 : with_variant<M, TabDP5, ALG, double>(records, stats, f);
}
template <class M, class Fn>
cudaError_t with_model_variant(int alg, int ode, int prec, bool records, bool stats, Fn&& f) {
    if (alg == PDMP_ALG_CHV) return with_alg_variant<M, PDMP_ALG_CHV>(ode, prec, records, stats, f);
    if constexpr (M::HAS_BOUND) return with_alg_variant<M, PDMP_ALG_REJECTION>(ode, prec, records, stats, f);
    return cudaErrorInvalidValue;
}

template <class M>
cudaError_t model_launch(int alg, int ode, int prec, bool records, bool stats, const KParams& P, int grid, size_t smem,
                         cudaStream_t st) {
    return with_model_variant<M>(alg, ode, prec, records, stats, [&](auto kern) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, BLOCK, smem, st>>>(P);
        return cudaGetLastError();
    });
}
template <class M>
cudaError_t model_occupancy(int alg, int ode, int prec, bool records, bool stats, size_t smem, int* out) {
    return with_model_variant<M>(alg, ode, prec, records, stats, [&](auto kern) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, kern, BLOCK, smem);
    });
}
template <class M, class XT>
cudaError_t compact_launch(bool co_resident, const KParams& P, const unsigned long long* offsets, double* time, double* xc,
                           XT* xd, double* rate, int sms, cudaStream_t st) {
    const unsigned long long want = (P.n_traj + 3) / 4;  // 4 warps per CTA, one trajectory per warp at a time
    const int grid = (int)std::min<unsigned long long>(want ? want : 1, (unsigned long long)sms * (co_resident ? 16 : PDMP_COMPACT_CTAS));
    if (co_resident) compact_kernel<M, XT, true><<<grid, 128, 0, st>>>(P, offsets, time, xc, xd, rate);
    else compact_kernel<M, XT, false><<<grid, 128, 0, st>>>(P, offsets, time, xc, xd, rate);
    return cudaGetLastError();
}
template <class M>
cudaError_t model_compact(const KParams& P, const unsigned long long* offsets, double* time, double* xc,
                          void* xd, double* rate, int xd_bytes, bool co_resident, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (xd_bytes == 1) return compact_launch<M, unsigned char>(co_resident, P, offsets, time, xc, (unsigned char*)xd, rate, sms, st);
    if (xd_bytes == 2) return compact_launch<M, short>(co_resident, P, offsets, time, xc, (short*)xd, rate, sms, st);
    if (xd_bytes == 4) return compact_launch<M, int>(co_resident, P, offsets, time, xc, (int*)xd, rate, sms, st);
    return compact_launch<M, long long>(co_resident, P, offsets, time, xc, (long long*)xd, rate, sms, st);
}

template <class M>
ModelEntry make_entry(int id) {
    ModelEntry e{};
    e.info.id = id;
    e.info.n_c = M::NC; e.info.n_d = M::ND; e.info.n_r = M::NR; e.info.n_parms = M::NP;
    e.info.has_bound = M::HAS_BOUND; e.info.has_delta = M::HAS_DELTA; e.info.zero_flow = M::ZERO_FLOW;
    for (int i = 0; M::NAME[i] && i < 31; ++i) e.info.name[i] = M::NAME[i];
    e.rec_words = RecLayout<M>::WORDS;
    e.sched_jump = SchedDefaults<M>::jump; e.sched_refill = SchedDefaults<M>::refill;
    e.const_rate_ok = ConstRateOk<M>::value;
    e.composite_ok = CompositeOk<M>::value;
    e.info.composite = CompositeOk<M>::value;
    e.launch = &model_launch<M>;
    e.occupancy = &model_occupancy<M>;
    e.compact = &model_compact<M>;
    return e;
}

}  // namespace pdmp
