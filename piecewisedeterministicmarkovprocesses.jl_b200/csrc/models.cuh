// models.cuh — the shipped models as CUDA device functors (one per reference example), and the
// documented hook for user models.
//
// A model is a struct with
//     static constexpr int  NC, ND, NR, NP;             continuous dim, discrete dim, reactions, #parms
//     static constexpr bool HAS_BOUND, HAS_DELTA, ZERO_FLOW;
//     static constexpr const char* NAME;
//     F(dx, x, xd, p, t)            vector field            (reference signature F(ẋ, xc, xd, parms, t))
//     rtot(x, xd, p, t)             total rate              (R(rate, xc, xd, parms, t, true)[1])
//     rates(r, x, xd, p, t)         all NR rates            (R(rate, xc, xd, parms, t, false))
//     bound(x, xd, p, t)            rate bound              (R(...)[2], Rejection only)
//     delta(x, xd, p, t, ev)        jump on xc after nu     (Delta(xc, xd, parms, t, ind_reaction), ev 1-based)
//   Every functor is a template on the real type R (double; float for the optional FP32 kernels):
//   state arrays and time are R, parameters stay `const double* p` (cast with R(p[i])), and
//   literals are written R(...) so that the FP32 kernels never promote to double.
//   optional fast path (HAS_CHV_FIELD): chv_field(dy, y, xd, p) = (F, 1) / rtot written out
//   algebraically, saving the reciprocal when the model allows it.
// x is indexable 0..NC-1 (the CHV kernel passes the extended state, time in slot NC, exactly
// like the reference passes `x` with the time slot to R and F: src/chvdiffeq.jl:6-16).
// xd is int32 in registers; the reference's Int64 is restored on output.
//
// USER MODELS: write such a struct here, add csrc/inst_<name>.cu (make_entry<YourDev>), list it in
// registry() at the top of abi.cu and rebuild (python -c "import __graft_entry__ as g; g.build()").
// See INTEGRATION.md.
#pragma once
#include "device_core.cuh"

namespace pdmp {

struct TcpDev {  // examples/tcp.jl:25-45
    static constexpr int NC = 1, ND = 2, NR = 2, NP = 1;
    static constexpr bool HAS_BOUND = true, HAS_DELTA = false, ZERO_FLOW = false, HAS_CHV_FIELD = true;
    static constexpr const char* NAME = "tcp";
    static constexpr int SCHED_JUMP = 16, SCHED_REFILL = 2;   // ~5.5 cheap RK attempts per jump
    static constexpr bool CONST_FLOW = true;                  // F = +-1 between jumps: sample times and the last bit are exact
    template <class R> PDMP_DEV static void F(R* dx, const R* x, const int* xd, const double* p, R t) {
        dx[0] = (xd[0] & 1) ? R(-1) : R(1);  // mod(xd[1], 2) == 0 ? 1 : -1
    }
    template <class R> PDMP_DEV static R rtot(const R* x, const int* xd, const double* p, R t) { return rcp_fast(x[0]); }
    template <class R> PDMP_DEV static void rates(R* r, const R* x, const int* xd, const double* p, R t) {
        r[0] = R(1) / x[0];
        r[1] = R(0);
    }
    template <class R> PDMP_DEV static R bound(const R* x, const int* xd, const double* p, R t) { return R(100); }
    template <class R> PDMP_DEV static void delta(R* x, int* xd, const double* p, double t, int ev) {}
    // (F, 1) / (1/x) = (±1, 1) * x : a sign vector times a scalar (rk_attempt_signed, device_core.cuh)
    static constexpr bool CHV_SIGNED_SCALAR = true;
    // the sign vector as a bit mask (bit i set: component i of the field is -g): (-1)^xd[1] on x, +1 on t
    PDMP_DEV static unsigned chv_negmask(const int* xd, const double* p) { return (unsigned)(xd[0] & 1); }
    template <class R> PDMP_DEV static R chv_scalar(const R* y, const int* xd, const double* p) { return y[0]; }
    template <class R> PDMP_DEV static void chv_field(R* dy, const R* y, const int* xd, const double* p) {
        const R sg = (R)(1 - 2 * (xd[0] & 1));  // hoisted out of the stage loop by the compiler
        dy[0] = sg * y[0];
        dy[1] = y[0];
    }
};

struct Tcp2dDev {  // examples/tcp2d.jl:4-25
    static constexpr int NC = 2, ND = 2, NR = 2, NP = 1;
    static constexpr bool HAS_BOUND = false, HAS_DELTA = false, ZERO_FLOW = false, HAS_CHV_FIELD = false;
    static constexpr const char* NAME = "tcp2d";
    static constexpr int SCHED_JUMP = 16, SCHED_REFILL = 4;   // short trajectories: batch the initialisations (sweep: profiles/r2_sched_sweep_c.txt)
    template <class R> PDMP_DEV static void F(R* dx, const R* x, const int* xd, const double* p, R t) {
        if ((xd[0] & 1) == 0) { dx[0] = R(1); dx[1] = R(-1) * x[1]; }
        else { dx[0] = R(-1) * x[0]; dx[1] = R(1); }
    }
    template <class R> PDMP_DEV static R rtot(const R* x, const int* xd, const double* p, R t) {
        return (x[0] + x[1]) + R(p[0]) * (R)xd[0] * x[0];
    }
    template <class R> PDMP_DEV static void rates(R* r, const R* x, const int* xd, const double* p, R t) {
        r[0] = x[0] + x[1];
        r[1] = R(p[0]) * (R)xd[0] * x[0];
    }
    template <class R> PDMP_DEV static R bound(const R* x, const int* xd, const double* p, R t) { return R(0); }
    template <class R> PDMP_DEV static void delta(R* x, int* xd, const double* p, double t, int ev) {}
    // exact flow of F over dt (either sign): one component moves at unit speed, the other decays exponentially
    static constexpr bool HAS_EXACT_FLOW = true;
    template <class R> PDMP_DEV static void flow(R* x, const int* xd, const double* p, R t, R dt) {
        const R e = exp(-dt);
        if ((xd[0] & 1) == 0) { x[0] += dt; x[1] *= e; }
        else { x[0] *= e; x[1] += dt; }
    }
};

struct MorrisLecarDev {  // examples/morris_lecar.jl:8-28
    static constexpr int NC = 1, ND = 4, NR = 4, NP = 15;
    static constexpr bool HAS_BOUND = false, HAS_DELTA = false, ZERO_FLOW = false, HAS_CHV_FIELD = true;
    static constexpr const char* NAME = "morris_lecar";
    static constexpr int SCHED_JUMP = 6, SCHED_REFILL = 2;    // ~2 expensive RK attempts (18 exp) per jump
    enum { v_Na, g_Na, v_K, g_K, v_L, g_L, I_app, gamma_na, k_na, beta_na, gamma_k, k_k, beta_k, M, N };
    template <class R> PDMP_DEV static void F(R* dx, const R* x, const int* xd, const double* p, R t) {
        dx[0] = div_fast((R)xd[1], R(p[N])) * (R(p[g_Na]) * (R(p[v_Na]) - x[0])) +
                div_fast((R)xd[3], R(p[M])) * (R(p[g_K]) * (R(p[v_K]) - x[0])) + (R(p[g_L]) * (R(p[v_L]) - x[0])) + R(p[I_app]);
    }
    template <class R> PDMP_DEV static void rates(R* r, const R* x, const int* xd, const double* p, R t) {
        r[0] = R(p[beta_na]) * exp(R(4) * R(p[gamma_na]) * x[0] + R(4) * R(p[k_na])) * (R)xd[0];
        r[1] = R(p[beta_na]) * (R)xd[1];
        r[2] = R(p[beta_k]) * exp(R(p[gamma_k]) * x[0] + R(p[k_k])) * (R)xd[2];
        r[3] = R(p[beta_k]) * exp(-R(p[gamma_k]) * x[0] - R(p[k_k])) * (R)xd[3];
    }
    template <class R> PDMP_DEV static R rtot(const R* x, const int* xd, const double* p, R t) {
        R r[4];
        rates(r, x, xd, p, t);
        return ((r[0] + r[1]) + r[2]) + r[3];
    }
    template <class R> PDMP_DEV static R bound(const R* x, const int* xd, const double* p, R t) { return R(0); }
    template <class R> PDMP_DEV static void delta(R* x, int* xd, const double* p, double t, int ev) {}
    // Jump cache (optional hook, HAS_JUMP_CACHE): a jump changes xd only, so the field at the post-jump state needs
    // the same two exponentials of v the rates at the pre-jump state have just evaluated — the same expressions, the
    // same values; the threshold section keeps them instead of calling exp twice more.
    static constexpr bool HAS_JUMP_CACHE = true;
    template <class R> struct JumpCache { R e_na, e_k; };
    template <class R> PDMP_DEV static void rates_cached(R* r, const R* x, const int* xd, const double* p, R t, JumpCache<R>& c) {
        c.e_na = exp(R(4) * R(p[gamma_na]) * x[0] + R(4) * R(p[k_na]));
        c.e_k = exp(R(p[gamma_k]) * x[0] + R(p[k_k]));
        r[0] = R(p[beta_na]) * c.e_na * (R)xd[0];
        r[1] = R(p[beta_na]) * (R)xd[1];
        r[2] = R(p[beta_k]) * c.e_k * (R)xd[2];
        r[3] = R(p[beta_k]) * exp(-R(p[gamma_k]) * x[0] - R(p[k_k])) * (R)xd[3];
    }
    template <class R> PDMP_DEV static void chv_field_cached(R* dy, const R* y, const int* xd, const double* p, const JumpCache<R>& c) {
        chv_field_from(dy, y, xd, p, c.e_na, c.e_k);
    }
    // The CHV field (F, 1) / Rtot with the total rate written for the integrator: the two potassium rates are
    // beta_k exp(+-(gamma_k v + k_k)), so ONE exponential serves both (2 exp per stage instead
    // of 3; the value differs from the three-exponential sum in the last bits only).  The rates used for the
    // categorical draw at a jump (`rates` above) keep the reference's three separate exponentials.
    template <class R> PDMP_DEV static void chv_field(R* dy, const R* y, const int* xd, const double* p) {
        const R v = y[0];
        chv_field_from(dy, y, xd, p, exp(R(4) * R(p[gamma_na]) * v + R(4) * R(p[k_na])), exp(R(p[gamma_k]) * v + R(p[k_k])));
    }
    template <class R> PDMP_DEV static void chv_field_from(R* dy, const R* y, const int* xd, const double* p, const R e_na, const R e_k) {
#ifdef PDMP_ML_TWO_RCP
        const R e_ki = rcp_fast(e_k);
        const R tot = ((R(p[beta_na]) * e_na * (R)xd[0] + R(p[beta_na]) * (R)xd[1]) + R(p[beta_k]) * e_k * (R)xd[2]) +
                      R(p[beta_k]) * e_ki * (R)xd[3];
        const R inv = rcp_fast(tot);
#elif defined(PDMP_ML_PLAIN_FIELD)
        // 1 / tot = e_k / (e_k tot): the term in 1 / e_k becomes a constant, one reciprocal instead of two on the
        // dependent path exp -> total rate -> field
        const R part = (R(p[beta_na]) * e_na * (R)xd[0] + R(p[beta_na]) * (R)xd[1]) + R(p[beta_k]) * e_k * (R)xd[2];
        const R inv = e_k * rcp_fast(part * e_k + R(p[beta_k]) * (R)xd[3]);
#else
        // 1 / tot = e_k / (e_k tot): the term in 1 / e_k becomes a constant, one reciprocal instead of two on the
        // dependent path exp -> total rate -> field.  Everything that depends on (xd, p) only is grouped into
        // coefficients — loop-invariant over the stages of an RK phase, the compiler hoists them — so that a stage
        // is 3 DFMA for the total rate and 1 for F = a - b v.
        const R c0 = R(p[beta_na]) * (R)xd[0], c1 = R(p[beta_na]) * (R)xd[1], c2 = R(p[beta_k]) * (R)xd[2], c3 = R(p[beta_k]) * (R)xd[3];
        const R inv = e_k * rcp_fast(fma(fma(c2, e_k, fma(c0, e_na, c1)), e_k, c3));
#endif
#if defined(PDMP_ML_TWO_RCP) || defined(PDMP_ML_PLAIN_FIELD)
        R f[1];
        F(f, y, xd, p, y[1]);
        dy[0] = f[0] * inv;
#else
        const R g_na = div_fast((R)xd[1], R(p[N])) * R(p[g_Na]), g_k = div_fast((R)xd[3], R(p[M])) * R(p[g_K]);
        const R b = (g_na + g_k) + R(p[g_L]);
        const R a = fma(g_na, R(p[v_Na]), fma(g_k, R(p[v_K]), fma(R(p[g_L]), R(p[v_L]), R(p[I_app]))));
        dy[0] = fma(-b, y[0], a) * inv;
#endif
        dy[1] = inv;
    }
};

template <bool REJ>
struct SirDevT {  // examples/sir-rejection.jl:3-17 (REJ) / examples/sir.jl:6-27 ; F = F_dummy
    static constexpr int NC = 1, ND = 4, NR = 3, NP = 2;
    static constexpr bool HAS_BOUND = REJ, HAS_DELTA = false, ZERO_FLOW = true, HAS_CHV_FIELD = false;
    static constexpr const char* NAME = REJ ? "sir_rejection" : "sir";
    static constexpr bool CONST_RATE_OK = true;    // rates depend on xd only
    template <class R> PDMP_DEV static void F(R* dx, const R* x, const int* xd, const double* p, R t) { dx[0] = R(0); }
    template <class R> PDMP_DEV static void rates(R* r, const R* x, const int* xd, const double* p, R t) {
        r[0] = R(p[0]) * (R)xd[0] * (R)xd[1];
        r[1] = R(p[1]) * (R)xd[1];
        r[2] = REJ ? R(p[0]) : R(0.01);
    }
    template <class R> PDMP_DEV static R rtot(const R* x, const int* xd, const double* p, R t) {
        R r[3];
        rates(r, x, xd, p, t);
        return r[0] + r[1] + r[2];
    }
    template <class R> PDMP_DEV static R bound(const R* x, const int* xd, const double* p, R t) { return R(p[0] + 3.5); }
    template <class R> PDMP_DEV static void delta(R* x, int* xd, const double* p, double t, int ev) {}
};
using SirRejDev = SirDevT<true>;
using SirChvDev = SirDevT<false>;

struct StiffDev {  // test/pdmpStiff.jl:71-90
    static constexpr int NC = 1, ND = 2, NR = 2, NP = 1;
    static constexpr bool HAS_BOUND = true, HAS_DELTA = false, ZERO_FLOW = false, HAS_CHV_FIELD = false;
    static constexpr const char* NAME = "pdmp_stiff";
    template <class R> PDMP_DEV static void F(R* dx, const R* x, const int* xd, const double* p, R t) {
        dx[0] = (xd[0] & 1) ? R(-3) * (x[0] * x[0]) : R(10) * x[0];
    }
    template <class R> PDMP_DEV static R rtot(const R* x, const int* xd, const double* p, R t) { return x[0] + R(p[0]); }
    template <class R> PDMP_DEV static void rates(R* r, const R* x, const int* xd, const double* p, R t) {
        r[0] = x[0];
        r[1] = R(p[0]);
    }
    template <class R> PDMP_DEV static R bound(const R* x, const int* xd, const double* p, R t) { return R(p[0] + 100.0); }
    template <class R> PDMP_DEV static void delta(R* x, int* xd, const double* p, double t, int ev) {}
};

struct StiffDeltaDev : StiffDev {  // test/pdmpStiff.jl:198-215: jumps written as a Delta! function, nu = 0
    static constexpr bool HAS_DELTA = true;
    static constexpr const char* NAME = "pdmp_stiff_delta";
    template <class R> PDMP_DEV static void delta(R* x, int* xd, const double* p, double t, int ev) {
        if (ev == 1) xd[0] += 1; else xd[1] -= 1;
    }
};

struct RatesCstDev {  // test/testRatesCst.jl:5-24: the flow of pdmp_stiff, rates that depend on xd only
    static constexpr int NC = 1, ND = 2, NR = 2, NP = 1;
    static constexpr bool HAS_BOUND = false, HAS_DELTA = false, ZERO_FLOW = false, HAS_CHV_FIELD = false;
    static constexpr bool CONST_RATE_OK = true;    // usable with ConstantRate (src/rate.jl:17-40)
    static constexpr const char* NAME = "rates_cst";
    template <class R> PDMP_DEV static void F(R* dx, const R* x, const int* xd, const double* p, R t) {
        dx[0] = (xd[0] & 1) ? R(-3) * (x[0] * x[0]) : R(10) * x[0];
    }
    template <class R> PDMP_DEV static R rtot(const R* x, const int* xd, const double* p, R t) {
        return (R(10) + (R)xd[0]) + R(p[0]);
    }
    template <class R> PDMP_DEV static void rates(R* r, const R* x, const int* xd, const double* p, R t) {
        r[0] = R(10) + (R)xd[0];
        r[1] = R(p[0]);
    }
    template <class R> PDMP_DEV static R bound(const R* x, const int* xd, const double* p, R t) { return R(0); }
    template <class R> PDMP_DEV static void delta(R* x, int* xd, const double* p, double t, int ev) {}
};

struct RatesCompositeDev {  // test/testRatesComposite.jl:5-45: the flow of pdmp_stiff; rate 1 = 10 + xc[1] varies
    static constexpr int NC = 1, ND = 2, NR = 2, NP = 1;   // between jumps (Rvar!), rate 2 = xd[2] does not (Rcst!)
    static constexpr bool HAS_BOUND = false, HAS_DELTA = false, ZERO_FLOW = false, HAS_CHV_FIELD = false;
    static constexpr bool COMPOSITE_OK = true;     // usable with CompositeRate (src/rate.jl:42-69)
    static constexpr const char* NAME = "rates_composite";
    template <class R> PDMP_DEV static void F(R* dx, const R* x, const int* xd, const double* p, R t) {
        dx[0] = (xd[0] & 1) ? R(-3) * (x[0] * x[0]) : R(10) * x[0];
    }
    // R!(rate, xc, xd, parms, t, true) = R(xc[1]) + xd[2]
    template <class R> PDMP_DEV static R rtot(const R* x, const int* xd, const double* p, R t) { return (R(10) + x[0]) + (R)xd[1]; }
    // Rcst!(..., true) = xd[2] and Rvar!(..., true) = R(xc[1]); CompositeRate returns out_cst .+ out_var
    template <class R> PDMP_DEV static R rtot_cst(const R* x, const int* xd, const double* p, R t) { return (R)xd[1]; }
    template <class R> PDMP_DEV static R rtot_var(const R* x, const int* xd, const double* p, R t) { return R(10) + x[0]; }
    template <class R> PDMP_DEV static void rates(R* r, const R* x, const int* xd, const double* p, R t) {
        r[0] = R(10) + x[0];
        r[1] = (R)xd[1];
    }
    template <class R> PDMP_DEV static R bound(const R* x, const int* xd, const double* p, R t) { return R(0); }
    template <class R> PDMP_DEV static void delta(R* x, int* xd, const double* p, double t, int ev) {}
};

struct TcpRejDev {  // examples/tcp_rejection.jl:5-25
    static constexpr int NC = 1, ND = 2, NR = 2, NP = 1;
    static constexpr bool HAS_BOUND = true, HAS_DELTA = false, ZERO_FLOW = false, HAS_CHV_FIELD = false;
    static constexpr const char* NAME = "tcp_rejection";
    template <class R> PDMP_DEV static void F(R* dx, const R* x, const int* xd, const double* p, R t) {
        dx[0] = (xd[0] & 1) ? R(-10) * x[0] : R(1);
    }
    template <class R> PDMP_DEV static R rtot(const R* x, const int* xd, const double* p, R t) {
        return R(1) / (R(1) + exp(-x[0]));
    }
    template <class R> PDMP_DEV static void rates(R* r, const R* x, const int* xd, const double* p, R t) {
        r[0] = R(1) / (R(1) + exp(-x[0]));
        r[1] = R(0);
    }
    template <class R> PDMP_DEV static R bound(const R* x, const int* xd, const double* p, R t) { return R(1); }
    template <class R> PDMP_DEV static void delta(R* x, int* xd, const double* p, double t, int ev) {}
};

struct ExplosionDev {  // examples/pdmpExplosion.jl:29-45, r = 10
    static constexpr int NC = 1, ND = 2, NR = 2, NP = 1;
    static constexpr bool HAS_BOUND = true, HAS_DELTA = false, ZERO_FLOW = false, HAS_CHV_FIELD = false;
    static constexpr const char* NAME = "pdmp_explosion";
    template <class R> PDMP_DEV static void F(R* dx, const R* x, const int* xd, const double* p, R t) {
        dx[0] = R(-10) * (R)(2 * (xd[0] & 1) - 1) * x[0];
    }
    template <class R> PDMP_DEV static R rtot(const R* x, const int* xd, const double* p, R t) { return x[0]; }
    template <class R> PDMP_DEV static void rates(R* r, const R* x, const int* xd, const double* p, R t) {
        r[0] = x[0];
        r[1] = R(0);
    }
    template <class R> PDMP_DEV static R bound(const R* x, const int* xd, const double* p, R t) { return R(40); }
    template <class R> PDMP_DEV static void delta(R* x, int* xd, const double* p, double t, int ev) {}
};

struct NeuronEvaDev {  // examples/pdmp_example_eva.jl:5-45 (Delta resets xc on reaction 2)
    static constexpr int NC = 1, ND = 2, NR = 3, NP = 1;
    static constexpr bool HAS_BOUND = true, HAS_DELTA = true, ZERO_FLOW = false, HAS_CHV_FIELD = false;
    static constexpr const char* NAME = "neuron_eva";
    template <class R> PDMP_DEV static void F(R* dx, const R* x, const int* xd, const double* p, R t) {
        dx[0] = -(x[0] - R(1.5));
    }
    template <class R> PDMP_DEV static R rtot(const R* x, const int* xd, const double* p, R t) {
        const R x2 = x[0] * x[0];
        return xd[0] == 0 ? x2 * x2 + R(p[0]) : R(1) + R(p[0]);
    }
    template <class R> PDMP_DEV static void rates(R* r, const R* x, const int* xd, const double* p, R t) {
        const R x2 = x[0] * x[0];
        if (xd[0] == 0) { r[0] = x2 * x2; r[1] = R(0); }
        else { r[0] = R(0); r[1] = R(1); }
        r[2] = R(p[0]);
    }
    template <class R> PDMP_DEV static R bound(const R* x, const int* xd, const double* p, R t) { return R(5); }
    template <class R> PDMP_DEV static void delta(R* x, int* xd, const double* p, double t, int ev) {
        if (ev == 2) x[0] = R(0);
    }
};

}  // namespace pdmp
