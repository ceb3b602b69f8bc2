// device_core.cuh — building blocks of the sm_100a ensemble kernels:
//   * Philox4x32-10 counter RNG + per-trajectory uniform stream (replaces the reference's bare
//     `rand()` calls on the global RNG: src/problem.jl:153, src/chvdiffeq.jl:71, src/utils.jl:47,
//     src/rejectiondiffeq.jl:35-36)
//   * Dormand-Prince 5(4) and Tsitouras 5(4) tableaus as compile-time constants
//   * one adaptive RK attempt with everything in registers (replaces OrdinaryDiffEq's
//     perform_step! called through SciMLBase.step!, reference call site src/chvdiffeq.jl:143)
//   * the PI step-size controller (OrdinaryDiffEq PIController defaults), powers in FP32 like
//     OrdinaryDiffEq's `fastpower`
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pdmp {

#define PDMP_DEV __device__ __forceinline__

// ---------------------------------------------------------------------------------------
// fast FP64 reciprocal: MUFU.RCP64H seed + 2 Newton steps (4 DFMA).  <= 2 ulp; used for vector-field scaling
// and error weights, never for quantities that decide a discrete outcome.  No division fallback: an
// `if (out of range) r = 1.0 / x` gets if-converted by the compiler, i.e. every call then pays for a full IEEE
// division as well (measured: +15 % on the Morris-Lecar kernel).  Every input the Newton steps cannot handle
// (0, inf, NaN, subnormals — flushed by the seed) turns them into NaN, and the bare seed is then the answer
// (inf, 0, NaN, inf).
// ---------------------------------------------------------------------------------------
PDMP_DEV double rcp_fast(double x) {
#ifdef PDMP_EXACT_DIV
    return 1.0 / x;
#else
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    double e = fma(-x, r0, 1.0);
    double r = fma(r0, e, r0);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r == r ? r : r0;
#endif
}

// a / b for operands of moderate size (no scaling for results near the overflow / underflow thresholds): Markstein's
// residual correction of a * (1 / b), the correctly rounded quotient except in rare last-bit cases.  Unlike the
// compiler's division it has no slow path — `0.0 / b` (an empty channel population over its size) sends the built-in
// sequence into its ~60-instruction subroutine every time.
PDMP_DEV double div_fast(double a, double b) {
    const double r = rcp_fast(b);
    const double q = a * r;
    return fma(fma(-q, b, a), r, q);
}
PDMP_DEV float div_fast(float a, float b) { return a / b; }

// error weights abstol + |u| reltol (positive, normal range) only feed the step controller, whose
// powers are FP32 anyway: the bare MUFU.RCP64H seed (>= 20 mantissa bits, PTX rcp.approx.ftz.f64)
// is enough — a relative error of 1e-6 in EEst moves an accept/reject decision only when EEst
// is within 1e-6 of 1
PDMP_DEV double rcp_weight(double x) {
#ifdef PDMP_EXACT_DIV
    return 1.0 / x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
#endif
}

// FP32 flavours (the optional PDMP_PREC_FP32 kernels)
PDMP_DEV float rcp_fast(float x) { return __frcp_rn(x); }
PDMP_DEV float rcp_weight(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// max(|a|, |b|) of the error weights as compare + select.  fmax's NaN rules are not needed here (a NaN
// stage makes the error estimate NaN whichever operand is picked, and the caller treats that as a
// failed trial step); for finite operands the value is the one fmax returns.
template <class R> PDMP_DEV R max_abs(R a, R b) {
    return fabs(fabs(a) > fabs(b) ? a : b);   // select the raw operand: |.| folds into the consumer as an operand modifier
}
// h with the sign of component i of a sign vector given as a bit mask (bit i set: c_i = -1)
PDMP_DEV double apply_sign(double h, unsigned negmask, int i) {
    return __hiloint2double(__double2hiint(h) ^ (int)(((negmask >> i) & 1u) << 31), __double2loint(h));
}
PDMP_DEV float apply_sign(float h, unsigned negmask, int i) {
    return __int_as_float(__float_as_int(h) ^ (int)(((negmask >> i) & 1u) << 31));
}

// ---------------------------------------------------------------------------------------
// -log(u) for a uniform u in (0, 1): the exponential variate of every threshold (reference
// `-log(rand())`, src/problem.jl:153, src/chvdiffeq.jl:71, src/rejectiondiffeq.jl:36).
// u = 2^e m with m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.172, by the
// odd series up to s^19 (truncation 2e-17); measured against libm: <= 4.5e-16 relative.  About 40
// instructions against ~85 for the library log: no denormal / special-case handling on the hot path
// (any u outside the normal range (0,1) takes the library call) and the coefficients are
// constant-bank operands of the DFMAs instead of pairs of 32-bit immediates.
// ---------------------------------------------------------------------------------------
static __constant__ double c_atanh[9] = {1.0 / 3, 1.0 / 5, 1.0 / 7, 1.0 / 9, 1.0 / 11, 1.0 / 13, 1.0 / 15, 1.0 / 17, 1.0 / 19};
static __constant__ double c_ln2[2] = {0.6931471805599453094, 2.319046813846299616e-17};
PDMP_DEV double neg_log_unit(double u) {
    int hi = __double2hiint(u);
    const int lo = __double2loint(u);
    if ((unsigned)(hi - 0x00100000) >= 0x3fe00000u) return -log(u);   // not a normal number in (0, 1)
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    if (hi >= 0x3ff6a09f) { hi -= 0x00100000; e += 1; }
    const double m = __hiloint2double(hi, lo);
    const double a = m - 1.0, b = m + 1.0;       // b in [1.70, 2.42)
    double rb;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rb) : "d"(b));
    double t = fma(-b, rb, 1.0);
    rb = fma(rb, t, rb);
    t = fma(-b, rb, 1.0);
    rb = fma(rb, t, rb);
    double s = a * rb;
    s = fma(rb, fma(-s, b, a), s);               // one residual correction of the quotient
    const double s2 = s * s;
    double p = c_atanh[8];
#pragma unroll
    for (int k = 7; k >= 0; --k) p = fma(p, s2, c_atanh[k]);
    const double lm = fma(s + s, s2 * p, s + s);  // 2 s + 2 s^3 P(s^2)
    const double ef = (double)e;
    return -fma(ef, c_ln2[0], fma(ef, c_ln2[1], lm));
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  counter = (gid_lo, gid_hi, block, 0),
// key = (seed_lo, seed_hi).  One call yields two 52-bit uniforms in (0,1): exactly one
// CHV jump's worth (pfsample draw + next exponential).
// ---------------------------------------------------------------------------------------
PDMP_DEV void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0;
        const uint32_t n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// the ten round keys (k0 + r W0, k1 + r W1) depend on the seed only: the host puts them in the kernel
// parameters, so that they reach the XORs as constant-bank operands instead of ~40 uniform-datapath
// instructions per call (Weyl increments + 32-bit immediates)
struct PhiloxKeys { uint32_t k[10][2]; };
inline PhiloxKeys philox_round_keys(uint64_t seed) {
    PhiloxKeys K;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { K.k[r][0] = k0; K.k[r][1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return K;
}
PDMP_DEV void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, const PhiloxKeys& K) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ K.k[r][0];
        const uint32_t n2 = h0 ^ c3 ^ K.k[r][1];
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    }
}
PDMP_DEV double u01_from_bits(uint32_t hi, uint32_t lo) {
    // (m + 1/2) 2^-52 with the 52-bit integer m = hi:lo >> 12, built without an integer-to-double
    // conversion: 1.m as a double in [1, 2), minus 1, plus 2^-53 — every step exact, the same value
    const uint32_t mh = hi >> 12, ml = (hi << 20) | (lo >> 12);
    return (__hiloint2double((int)(0x3ff00000u | mh), (int)ml) - 1.0) + 1.1102230246251565e-16;
}

struct DrawStream {
    uint32_t idx;     // draws consumed
    double cache;     // second half of the last Philox block
    PDMP_DEV void init() { idx = 0; cache = 0.0; }
    // REPLAY: reads the host-provided stream; PHILOX: draw k = block k/2, half k%2
    PDMP_DEV double next(int rng_mode, const double* __restrict__ replay, uint64_t replay_len, uint64_t gid,
                         const PhiloxKeys& keys, bool& exhausted) {
        const uint32_t k = idx++;
        if (rng_mode == 1) {
            if (k >= replay_len) { exhausted = true; return 0.5; }
            return __ldg(replay + k);
        }
        if (k & 1u) return cache;
        uint32_t c0 = (uint32_t)gid, c1 = (uint32_t)(gid >> 32), c2 = k >> 1, c3 = 0u;
        philox4x32_10(c0, c1, c2, c3, keys);
        cache = u01_from_bits(c2, c3);
        return u01_from_bits(c0, c1);
    }
    // The next TWO draws with exactly one Philox call whatever the parity of idx (the same stream as two
    // next() calls): draws k, k+1 need block (k+1)/2 only — for even k both halves of it, for odd k the
    // cached half of the previous block and the first half of it.  Lanes whose idx parities differ
    // (Rejection: 2 or 3 draws per candidate) therefore stay converged through the Philox rounds.
    PDMP_DEV void next2(int rng_mode, const double* __restrict__ replay, uint64_t replay_len, uint64_t gid,
                        const PhiloxKeys& keys, bool& exhausted, double& u0, double& u1) {
        const uint32_t k = idx;
        idx += 2;
        if (rng_mode == 1) {
            if (k + 1 >= replay_len) { exhausted = true; u0 = u1 = 0.5; return; }
            u0 = __ldg(replay + k); u1 = __ldg(replay + k + 1);
            return;
        }
        uint32_t c0 = (uint32_t)gid, c1 = (uint32_t)(gid >> 32), c2 = (k + 1u) >> 1, c3 = 0u;
        philox4x32_10(c0, c1, c2, c3, keys);
        const double a = u01_from_bits(c0, c1), b = u01_from_bits(c2, c3);
        if (k & 1u) { u0 = cache; u1 = a; cache = b; }
        else { u0 = a; u1 = b; }
    }
    // The same for a stream position that is KNOWN to be odd (a CHV trajectory has consumed 1 + 2 j draws at
    // its j-th threshold): the cached half, then the first half of a new block — no parity selects.
    PDMP_DEV void next2_odd(int rng_mode, const double* __restrict__ replay, uint64_t replay_len, uint64_t gid,
                            const PhiloxKeys& keys, bool& exhausted, double& u0, double& u1) {
        const uint32_t k = idx;
        idx += 2;
        if (rng_mode == 1) {
            if (k + 1 >= replay_len) { exhausted = true; u0 = u1 = 0.5; return; }
            u0 = __ldg(replay + k); u1 = __ldg(replay + k + 1);
            return;
        }
        uint32_t c0 = (uint32_t)gid, c1 = (uint32_t)(gid >> 32), c2 = (k + 1u) >> 1, c3 = 0u;
        philox4x32_10(c0, c1, c2, c3, keys);
        u0 = cache;
        u1 = u01_from_bits(c0, c1);
        cache = u01_from_bits(c2, c3);
    }
};

// ---------------------------------------------------------------------------------------
// Tableaus.  a(s, j): coefficient of k_{j+1} in stage s+2 (s = 0..5; s = 5 is the b row),
// bt(j): b - bhat.  constexpr functions so that fully unrolled loops fold the zeros away.
// ---------------------------------------------------------------------------------------
// Coefficient VALUES are read from __constant__ memory (they become c[bank][offset] operands of
// DFMA/DMUL: no instruction is spent materialising 64-bit immediates); the constexpr copies
// only drive compile-time structure (which coefficients are zero).
static __constant__ double c_dp5_a[6][6] = {
    {1.0 / 5, 0, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
    {35.0 / 384, 0.0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
static __constant__ double c_dp5_bt[7] = {71.0 / 57600, 0.0, -71.0 / 16695, 71.0 / 1920, -17253.0 / 339200,
                                          22.0 / 525, -1.0 / 40};
static __constant__ double c_dp5_c[7] = {0.0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};
static __constant__ double c_ts5_a[6][6] = {
    {0.161, 0, 0, 0, 0, 0},
    {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};
static __constant__ double c_ts5_bt[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
                                          -0.1447110071732629,     0.5823571654525552,     -0.45808210592918697,
                                          1.0 / 66.0};
static __constant__ double c_ts5_c[7] = {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0};

// FP32 copies of the tableaus (rounded once, at compile time)
static __constant__ float c_dp5_af[6][6] = {
    {(float)(1.0 / 5), (float)(0), (float)(0), (float)(0), (float)(0), (float)(0)},
    {(float)(3.0 / 40), (float)(9.0 / 40), (float)(0), (float)(0), (float)(0), (float)(0)},
    {(float)(44.0 / 45), (float)(-56.0 / 15), (float)(32.0 / 9), (float)(0), (float)(0), (float)(0)},
    {(float)(19372.0 / 6561), (float)(-25360.0 / 2187), (float)(64448.0 / 6561), (float)(-212.0 / 729), (float)(0), (float)(0)},
    {(float)(9017.0 / 3168), (float)(-355.0 / 33), (float)(46732.0 / 5247), (float)(49.0 / 176), (float)(-5103.0 / 18656), (float)(0)},
    {(float)(35.0 / 384), (float)(0.0), (float)(500.0 / 1113), (float)(125.0 / 192), (float)(-2187.0 / 6784), (float)(11.0 / 84)}};
static __constant__ float c_dp5_btf[7] = {(float)(71.0 / 57600), (float)(0.0), (float)(-71.0 / 16695), (float)(71.0 / 1920), (float)(-17253.0 / 339200),
                                          (float)(22.0 / 525), (float)(-1.0 / 40)};
static __constant__ float c_dp5_cf[7] = {(float)(0.0), (float)(1.0 / 5), (float)(3.0 / 10), (float)(4.0 / 5), (float)(8.0 / 9), (float)(1.0), (float)(1.0)};
static __constant__ float c_ts5_af[6][6] = {
    {(float)(0.161), (float)(0), (float)(0), (float)(0), (float)(0), (float)(0)},
    {(float)(-0.008480655492356989), (float)(0.335480655492357), (float)(0), (float)(0), (float)(0), (float)(0)},
    {(float)(2.8971530571054935), (float)(-6.359448489975075), (float)(4.3622954328695815), (float)(0), (float)(0), (float)(0)},
    {(float)(5.325864828439257), (float)(-11.748883564062828), (float)(7.4955393428898365), (float)(-0.09249506636175525), (float)(0), (float)(0)},
    {(float)(5.86145544294642), (float)(-12.92096931784711), (float)(8.159367898576159), (float)(-0.071584973281401), (float)(-0.028269050394068383), (float)(0)},
    {(float)(0.09646076681806523), (float)(0.01), (float)(0.4798896504144996), (float)(1.379008574103742), (float)(-3.290069515436081), (float)(2.324710524099774)}};
static __constant__ float c_ts5_btf[7] = {(float)(-0.00178001105222577714), (float)(-0.0008164344596567469), (float)(0.007880878010261995),
                                          (float)(-0.1447110071732629),     (float)(0.5823571654525552),     (float)(-0.45808210592918697),
                                          (float)(1.0 / 66.0)};
static __constant__ float c_ts5_cf[7] = {(float)(0.0), (float)(0.161), (float)(0.327), (float)(0.9), (float)(0.9800255409045097), (float)(1.0), (float)(1.0)};

struct TabDP5 {
    static constexpr int ID = 0;
    template <class R> PDMP_DEV static R A(int s, int j) {
        if constexpr (sizeof(R) == 4) return c_dp5_af[s][j]; else return c_dp5_a[s][j];
    }
    template <class R> PDMP_DEV static R BT(int j) {
        if constexpr (sizeof(R) == 4) return c_dp5_btf[j]; else return c_dp5_bt[j];
    }
    template <class R> PDMP_DEV static R C(int i) {
        if constexpr (sizeof(R) == 4) return c_dp5_cf[i]; else return c_dp5_c[i];
    }
    __host__ __device__ static constexpr double c(int i) {
        constexpr double C[7] = {0.0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};
        return C[i];
    }
    __host__ __device__ static constexpr double a(int s, int j) {
        constexpr double A[6][6] = {
            {1.0 / 5, 0, 0, 0, 0, 0},
            {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
            {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
            {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
            {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
            {35.0 / 384, 0.0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
        return A[s][j];
    }
    __host__ __device__ static constexpr double bt(int j) {
        constexpr double B[7] = {71.0 / 57600, 0.0, -71.0 / 16695, 71.0 / 1920, -17253.0 / 339200, 22.0 / 525,
                                 -1.0 / 40};
        return B[j];
    }
};

struct TabTsit5 {
    static constexpr int ID = 1;
    template <class R> PDMP_DEV static R A(int s, int j) {
        if constexpr (sizeof(R) == 4) return c_ts5_af[s][j]; else return c_ts5_a[s][j];
    }
    template <class R> PDMP_DEV static R BT(int j) {
        if constexpr (sizeof(R) == 4) return c_ts5_btf[j]; else return c_ts5_bt[j];
    }
    template <class R> PDMP_DEV static R C(int i) {
        if constexpr (sizeof(R) == 4) return c_ts5_cf[i]; else return c_ts5_c[i];
    }
    __host__ __device__ static constexpr double c(int i) {
        constexpr double C[7] = {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0};
        return C[i];
    }
    __host__ __device__ static constexpr double a(int s, int j) {
        constexpr double A[6][6] = {
            {0.161, 0, 0, 0, 0, 0},
            {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
            {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
            {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
            {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383,
             0},
            {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
             2.324710524099774}};
        return A[s][j];
    }
    __host__ __device__ static constexpr double bt(int j) {
        constexpr double B[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
                                 -0.1447110071732629, 0.5823571654525552, -0.45808210592918697, 1.0 / 66.0};
        return B[j];
    }
};

// ---------------------------------------------------------------------------------------
// One RK attempt of size h from (u, k1 = f(u)) : fills unew, k7 = f(unew) (FSAL) and returns
// e2 = sum_i (err_i / (abstol + max(|u_i|,|unew_i|) reltol))^2  (EEst^2 * D).
// All arrays are compile-time indexed: they live in registers.
// ---------------------------------------------------------------------------------------
template <class Tab, int D, class R, class RHS>
PDMP_DEV R rk_attempt(RHS&& f, const R (&u)[D], const R (&k1)[D], R t, R h, R (&unew)[D], R (&k7)[D], R reltol,
                      R abstol, R& inc_last) {
    R k[7][D];
#pragma unroll
    for (int i = 0; i < D; ++i) k[0][i] = k1[i];
#pragma unroll
    for (int s = 1; s < 7; ++s) {
        R us[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            R acc = Tab::template A<R>(s - 1, 0) * k[0][i];
#pragma unroll
            for (int j = 1; j < s; ++j)
                if (Tab::a(s - 1, j) != 0.0) acc = fma(Tab::template A<R>(s - 1, j), k[j][i], acc);
            us[i] = fma(h, acc, u[i]);
            if (s == 6 && i == D - 1) inc_last = h * acc;   // increment of the last component (FP32 time shadow)
        }
        if (s == 6) {
#pragma unroll
            for (int i = 0; i < D; ++i) unew[i] = us[i];
        }
        f(us, k[s], fma(Tab::template C<R>(s), h, t));
    }
    R e2 = R(0);
#pragma unroll
    for (int i = 0; i < D; ++i) {
        R acc = Tab::template BT<R>(0) * k[0][i];
#pragma unroll
        for (int j = 1; j < 7; ++j)
            if (Tab::bt(j) != 0.0) acc = fma(Tab::template BT<R>(j), k[j][i], acc);
        const R sc = fma(max_abs(u[i], unew[i]), reltol, abstol);
        const R r = (h * acc) * rcp_weight(sc);
        e2 = fma(r, r, e2);
        k7[i] = k[6][i];
    }
    return e2;
}

// The same attempt for a field of the form  f(u) = c * g(u)  with a SIGN vector c (c_i = +-1, constant
// over the step) and a scalar g: every stage slope is c times a scalar, so ONE accumulation chain
// serves all components (k_j[i] = c_i g_j exactly, hence sum_j a_sj k_j[i] = c_i sum_j a_sj g_j bit for
// bit: the results are identical to rk_attempt, with about half the FMAs and a third of the stage
// registers).  TCP in CHV variables is such a field: (F, 1) / Rtot = (+-1, 1) * x.
// The sign vector arrives as a bit mask (bit i set: c_i = -1) and is applied to h by flipping its sign bit.
template <class Tab, int D, class R, class G>
PDMP_DEV R rk_attempt_signed(G&& g, unsigned negmask, const R (&u)[D], R g1, R t, R h, R (&unew)[D], R& g7, R reltol,
                             R abstol, R& inc_last) {
    R gk[7], hc[D];
    gk[0] = g1;
#pragma unroll
    for (int i = 0; i < D; ++i) hc[i] = apply_sign(h, negmask, i);
#pragma unroll
    for (int s = 1; s < 7; ++s) {
        R acc = Tab::template A<R>(s - 1, 0) * gk[0];
#pragma unroll
        for (int j = 1; j < s; ++j)
            if (Tab::a(s - 1, j) != 0.0) acc = fma(Tab::template A<R>(s - 1, j), gk[j], acc);
        R us[D];
#pragma unroll
        for (int i = 0; i < D; ++i) us[i] = fma(hc[i], acc, u[i]);
        if (s == 6) {
            inc_last = h * acc;
#pragma unroll
            for (int i = 0; i < D; ++i) unew[i] = us[i];
        }
        gk[s] = g(us);
    }
    R acc = Tab::template BT<R>(0) * gk[0];
#pragma unroll
    for (int j = 1; j < 7; ++j)
        if (Tab::bt(j) != 0.0) acc = fma(Tab::template BT<R>(j), gk[j], acc);
    const R he = h * acc;
    R e2 = R(0);
#pragma unroll
    for (int i = 0; i < D; ++i) {
        const R sc = fma(max_abs(u[i], unew[i]), reltol, abstol);
        const R r = he * rcp_weight(sc);
        e2 = fma(r, r, e2);
    }
    g7 = gk[6];
    return e2;
}

// ---------------------------------------------------------------------------------------
// PI controller (OrdinaryDiffEq PIController defaults for order-5 pairs: beta1 = 7/50,
// beta2 = 2/25, gamma = 9/10, qmin = 1/5, qmax = 10, qoldinit = 1e-4).  The powers are
// evaluated in FP32 (MUFU lg2/ex2), as OrdinaryDiffEq's fastpower does.
//   accept : h_next = h / q,  qold = max(EEst, 1e-4)
//   reject : h_next = h / min(1/qmin, q11 / gamma)
// Works on EEst^2 (= e2 / D) to skip the square root.
// ---------------------------------------------------------------------------------------
PDMP_DEV float lg2_approx(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
PDMP_DEV float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

struct PIController {
    // log2 of qold^2, qold = max(EEst_prev, 1e-4): the controller works in the log2 domain on
    // EEst^2, so one MUFU.LG2 and one MUFU.EX2 per attempt and no square root
    float lq2;
    PDMP_DEV void init() { lq2 = -26.575424759f; }  // log2(1e-8)
    // returns the factor f such that h_next = h * f
    //   accept : f = 1/q,  q = clamp(EEst^b1 / qold^b2 / gamma, 1/qmax, 1/qmin)
    //   reject : f = 1/min(1/qmin, EEst^b1 / gamma)
    // The same two formulas on e2 = D EEst^2 straight from the attempt (log2 D is subtracted in the log
    // domain), as separate entry points: accepted steps are the hot path, rejections are rare.
    template <int D> PDMP_DEV float accept(float e2) {
        const float l = lg2_approx(e2) - (D == 1 ? 0.0f : D == 2 ? 1.0f : D == 3 ? 1.58496250072f : D == 4 ? 2.0f : __log2f((float)D));
        float lq = 0.07f * l - 0.04f * lq2 + 0.15200309344f;
        lq = fminf(fmaxf(lq, -3.32192809489f), 2.32192809489f);
        lq2 = fmaxf(l, -26.575424759f);
        return ex2_approx(-lq);
    }
    template <int D> PDMP_DEV float reject(float e2) const {
        const float l = lg2_approx(e2) - (D == 1 ? 0.0f : D == 2 ? 1.0f : D == 3 ? 1.58496250072f : D == 4 ? 2.0f : __log2f((float)D));
        return ex2_approx(-fminf(0.07f * l + 0.15200309344f, 2.32192809489f));
    }
    PDMP_DEV float step(float EEst2, bool accept) {
        const float l = lg2_approx(EEst2);            // log2(EEst^2); -inf for EEst = 0
        const float lq11 = 0.07f * l;                        // log2(EEst^beta1), beta1 = 7/50
        const float lg = 0.15200309344f;                     // -log2(gamma), gamma = 9/10
        float lq;
        if (accept) {
            lq = lq11 - 0.04f * lq2 + lg;                    // beta2 = 2/25
            lq = fminf(fmaxf(lq, -3.32192809489f), 2.32192809489f);   // [log2(1/10), log2(5)]
            lq2 = fmaxf(l, -26.575424759f);                  // qold = max(EEst, 1e-4)
        } else {
            lq = fminf(lq11 + lg, 2.32192809489f);
        }
        return ex2_approx(-lq);
    }
};

}  // namespace pdmp
