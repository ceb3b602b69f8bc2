/*
 * pdmp_b200.h — C ABI of libpdmp_b200.so, the B200-native ensemble engine for the
 * data-parallel hot path of PiecewiseDeterministicMarkovProcesses.jl (reference v0.0.11).
 *
 * The reference has no FFI: its extension point is Julia dispatch on an algorithm type plus a
 * `SciMLBase.solve(problem::PDMPProblem, algo; kwargs...)` method (reference
 * src/chvdiffeq.jl:211-237, src/rejectiondiffeq.jl:155-176).  This header is what a
 * `CHVGPU` / `RejectionGPU` solve method binds with `ccall`.  Plain
 * pointers and sizes only; no C++ or torch types cross the boundary.
 *
 * Conventions
 *  - every entry point returns 0 on success, a PDMP_ERR_* code otherwise; the message is
 *    available from pdmp_last_error() (thread-local).  No exception crosses the ABI.
 *  - per-trajectory problems (reference: @assert src/chvdiffeq.jl:145,
 *    src/rejectiondiffeq.jl:33, @warn :121-124) are reported in result.status[], never as
 *    a return code.
 *  - inputs are borrowed for the duration of the call; outputs are owned by the library
 *    and released with pdmp_result_free / pdmp_ensemble_destroy.
 *  - all structs contain only 8-byte and 4-byte scalars and pointers, 8-byte fields first
 *    inside each group, so that the natural C layout has no hidden padding surprises for
 *    ctypes / Julia `struct` mirrors.
 */
#ifndef PDMP_B200_H
#define PDMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDMP_ABI_VERSION 8

#define PDMP_STAT_BATCHES 64

/* ---- enumerations ------------------------------------------------------------------ */

/* algorithm: reference `CHV(ode)` (src/chvdiffeq.jl:2-4) / `Rejection(ode)`
 * (src/rejectiondiffeq.jl:1-3) */
/* REJECTION_EXACT: thinning with a user-supplied EXACT flow instead of an ODE solve, reference
 * `RejectionExact()` (src/rejection.jl:1, :3-91, :113-116); for models registered with
 * exact_flow = 1 (many continuous variables: one WARP per trajectory, not one thread). */
enum { PDMP_ALG_CHV = 0, PDMP_ALG_REJECTION = 1, PDMP_ALG_REJECTION_EXACT = 2 };

/* explicit adaptive 5(4) pair used for the flow; TSIT5 is the reference's `Tsit5()`,
 * DP5 the Dormand-Prince pair named by the north star. Same controller for both. */
enum { PDMP_ODE_DP5 = 0, PDMP_ODE_TSIT5 = 1 };

enum { PDMP_RNG_PHILOX = 0, PDMP_RNG_REPLAY = 1 };

/* FP32: state and Runge-Kutta stage arithmetic in float; the integration variable, thresholds,
 * step size and jump times stay double (CHV: the time slot is refreshed from a double shadow
 * after every accepted step).  Meant for reltol >= 1e-6; models whose state leaves the float
 * range (TCP on long horizons) flag the affected trajectories NO_PROGRESS / NAN. */
enum { PDMP_PREC_FP64 = 0, PDMP_PREC_FP32 = 1 };

/* per-trajectory status (result.status[i]) */
enum {
    PDMP_ST_OK = 0,              /* reached tf                                             */
    PDMP_ST_HIT_NJUMPS = 1,      /* stopped by the n_jumps cap (reference loop test
                                    src/chvdiffeq.jl:141)                                   */
    PDMP_ST_NO_PROGRESS = 2,     /* jump time did not advance (reference @assert
                                    src/chvdiffeq.jl:145 / @warn rejectiondiffeq.jl:121)    */
    PDMP_ST_NONPOSITIVE_RATE = 3,/* total rate <= 0 or not finite in the CHV field          */
    PDMP_ST_BOUND_VIOLATED = 4,  /* Rejection: total rate >= bound (reference @assert
                                    src/rejectiondiffeq.jl:33)                              */
    PDMP_ST_NAN = 5,             /* non-finite state / error estimate                       */
    PDMP_ST_MAX_ATTEMPTS = 6,    /* RK attempt budget exhausted (OrdinaryDiffEq maxiters)   */
    PDMP_ST_OUTPUT_OVERFLOW = 7, /* record pool exhausted: trajectory truncated             */
    PDMP_ST_REPLAY_EXHAUSTED = 8 /* replay stream shorter than the draws consumed           */
};

/* return codes */
enum {
    PDMP_OK = 0,
    PDMP_ERR_INVALID = 1,     /* bad argument                                              */
    PDMP_ERR_MODEL = 2,       /* unknown model / shape mismatch with the device functor    */
    PDMP_ERR_CUDA = 3,        /* CUDA runtime error (message in pdmp_last_error)           */
    PDMP_ERR_NO_DEVICE = 4,   /* no usable GPU: there is NO CPU fallback                   */
    PDMP_ERR_CAPACITY = 5,    /* record pool too small; result.records_needed tells how
                                 much the run wanted                                       */
    PDMP_ERR_UNSUPPORTED = 6
};

/* ---- model registry ---------------------------------------------------------------- */

/* A model is the compiled-in device functor triple (F, R, Delta) of one reference example;
 * Julia closures cannot cross to the device (reference signatures: src/problem.jl:101-131). */
typedef struct pdmp_model_info {
    int32_t id;
    int32_t n_c;        /* continuous dimension                                           */
    int32_t n_d;        /* discrete dimension  (= columns of nu)                          */
    int32_t n_r;        /* number of reactions (= rows of nu)                             */
    int32_t n_parms;    /* length of the flat parameter vector the functor reads          */
    int32_t has_bound;  /* R returns (total, bound): usable with PDMP_ALG_REJECTION        */
    int32_t has_delta;  /* model has a Delta(xc, xd, p, t, ev) acting after nu            */
    int32_t zero_flow;  /* F == F_dummy (reference src/problem.jl:4-7)                    */
    int32_t exact_flow; /* the functor is an exact flow Phi: PDMP_ALG_REJECTION_EXACT only   */
    int32_t composite;  /* R is split into a constant and a variable part: rate_kind 2 usable  */
    char name[32];
} pdmp_model_info;

/* ---- problem / options ------------------------------------------------------------- */

/* Mirrors PDMPProblem(F,R,[Delta],nu,xc0,xd0,parms,tspan) (reference src/problem.jl:168-193)
 * for one shard of an ensemble. */
typedef struct pdmp_problem_desc {
    uint64_t n_traj;          /* trajectories in THIS call (shard)                         */
    uint64_t traj_id_offset;  /* global id of local trajectory 0: Philox counters use the
                                 global id so results do not depend on the sharding        */
    const double* xc0;        /* [n_c] (broadcast) or [n_traj][n_c]                        */
    const int64_t* xd0;       /* [n_d] (broadcast) or [n_traj][n_d]                        */
    const int64_t* nu;        /* [n_r][n_d] row-major, or column-major if nu_col_major     */
    const double* parms;      /* [n_parms]                                                 */
    double t0, tf;            /* tspan                                                     */
    int32_t model_id;
    int32_t xc0_per_traj;     /* 0: broadcast, 1: per trajectory                           */
    int32_t xd0_per_traj;
    int32_t nu_col_major;     /* Julia passes column-major matrices                        */
    int32_t n_parms;
    int32_t rate_kind;        /* 0: VariableRate (reference default, src/problem.jl:55)
                                 1: ConstantRate (src/rate.jl:17-40): the total rate is constant
                                    between jumps; the CHV field re-uses the value computed at the
                                    post-jump state instead of re-evaluating R at every stage.
                                    Only for models whose rates depend on xd alone
                                    (PDMP_ERR_UNSUPPORTED otherwise: it would change results)
                                 2: CompositeRate (src/rate.jl:42-69): a ConstantRate part, cached from
                                    jump to jump, plus a VariableRate part evaluated at every stage.
                                    Only for models registered with such a split (composite = 1 in
                                    pdmp_model_info); results identical to rate_kind 0
                                    (test/testRatesComposite.jl:95-105)                         */
} pdmp_problem_desc;

/* Mirrors the keyword arguments of the reference `solve` (src/chvdiffeq.jl:187-219,
 * src/rejectiondiffeq.jl:155-163) plus the ensemble additions. */
typedef struct pdmp_solve_opts {
    double reltol, abstol;        /* reference defaults 1e-7 / 1e-9                        */
    int64_t n_jumps;              /* reference meaning: counts the initial point, the loop
                                     ends after n_jumps-1 thresholds (src/chvdiffeq.jl:136,141);
                                     must be >= 1 and finite                               */
    int64_t max_attempts;         /* RK attempt budget per trajectory; <=0: 1000000        */
    uint64_t seed;                /* Philox key                                            */
    const double* replay_u;       /* REPLAY: uniforms, trajectory i reads
                                     replay_u[i*replay_stride + k], k = 0,1,...            */
    uint64_t replay_stride;
    const double* sample_times;   /* [n_sample_times], increasing, within (t0, tf]         */
    const int32_t* hist_comp;     /* [hist_n] component index: 0..n_c-1 continuous,
                                     n_c..n_c+n_d-1 discrete                               */
    const double* hist_lo;        /* [hist_n]                                              */
    const double* hist_hi;        /* [hist_n]                                              */
    uint64_t max_records;         /* capacity of the record pool for this call; 0: sized by
                                     the library from free memory                          */
    int32_t algorithm;            /* PDMP_ALG_*                                            */
    int32_t ode;                  /* PDMP_ODE_*                                            */
    int32_t precision;            /* PDMP_PREC_*                                           */
    int32_t rng_mode;             /* PDMP_RNG_*                                            */
    int32_t save_pre;             /* save_positions[1]                                     */
    int32_t save_post;            /* save_positions[2]                                     */
    int32_t no_records;           /* 1: statistics only, not even the initial/final points */
    int32_t n_sample_times;
    int32_t hist_n;
    int32_t hist_bins;
    int32_t device;               /* CUDA device ordinal for this call                     */
    int32_t tstop_policy;         /* 0: after a step clipped by a threshold, propose
                                        h_clipped/q (OrdinaryDiffEq behaviour)
                                     1: additionally never propose less than the unclipped
                                        proposal that was pending before the clip          */
    int32_t ref_quirks;           /* 1: the final point at t = tf exactly as the reference computes it
                                     (src/chvdiffeq.jl:160-168, src/rejectiondiffeq.jl:142-150): a fresh
                                     plain-F solve at OrdinaryDiffEq's DEFAULT tolerances 1e-3 / 1e-6 with
                                     its automatic initial step, from caract.xc (the state at the last
                                     executed jump) at time tprev (the previous threshold / candidate).
                                     0: the same flow at the caller's tolerances, consistent anchor   */
    int32_t xd_bytes;             /* width of the saved discrete states in result.xd:
                                     0 or 8: int64 (the reference's Int64 layout)
                                     4: int32 — always exact (the kernel holds xd as int32)
                                     2: int16 — the caller asserts |xd| < 32768 for the whole
                                        run (e.g. |xd0| + n_jumps max|nu| without a Delta);
                                        the host bindings check this bound before asking.
                                     1: event transport — ONE byte per saved point: the index (1-based) of
                                        the reaction that changed xd since the trajectory's previous saved
                                        point, 0 for none (initial / final / pre-jump points).  xd is rebuilt
                                        by the bindings as xd0 + a prefix sum of nu rows.  Only with post-jump
                                        saving (save_post), for models without a Delta and
                                        |xd0| + (n_jumps - 1) max|nu| < 2^23 (checked by
                                        the library: PDMP_ERR_UNSUPPORTED otherwise).
                                     Narrow transport cuts the device->host bytes per saved
                                     point; the bindings widen per trajectory on access.    */
    int32_t save_c_count;         /* REJECTION_EXACT: save only the first save_c_count
                                     continuous / save_d_count discrete components of every  */
    int32_t save_d_count;         /* point (0: all — the reference saves all of them,
                                     src/rejection.jl:32); result.n_c / n_d report the counts */
    /* ---- ABI v8 ---- */
    const int32_t* devices;       /* [n_devices] CUDA ordinals: the ensemble is sharded over them by
                                     contiguous global id range, one host thread per device, statistics
                                     summed in the library; NULL / n_devices <= 1: `device` alone.  An
                                     ordinal may be listed more than once (several shards on one GPU). */
    uint64_t save_c_mask;         /* ind_save_c / ind_save_d of the reference (src/chv.jl:54-62, doc    */
    uint64_t save_d_mask;         /* src/chvdiffeq.jl:200-201) as bit masks: bit j set = component j is
                                     saved; 0 = all.  result.n_c / n_d are the saved counts and the
                                     columns of result.xc / xd follow increasing j                     */
    int32_t n_devices;
    int32_t save_rate;            /* 1: result.rate[k] = total rate at the pre-jump state of the jump
                                     that produced saved point k (reference save_rate / rate_hist,
                                     src/chvdiffeq.jl:47,54, src/rejectiondiffeq.jl:134); NaN for the
                                     initial and the final point                                       */
} pdmp_solve_opts;

/* ---- result ------------------------------------------------------------------------ */

/* Ensemble analogue of PDMPResult (reference src/utils.jl:66-79): CSR-ragged saved points.
 * Trajectory i owns records [offsets[i], offsets[i+1]).  time[k], xc[k*n_c + j],
 * xd[k*n_d + j]: as a Julia column-major matrix that is xc[j, k], the reference's layout. */
typedef struct pdmp_result {
    uint64_t n_traj;
    uint64_t total_records;
    uint64_t records_needed;   /* >= total_records; larger when PDMP_ERR_CAPACITY          */
    uint64_t* offsets;         /* [n_traj+1]                                               */
    double* time;              /* [total_records]                                          */
    double* xc;                /* [total_records][n_c]                                     */
    void* xd;                  /* [total_records][n_d], elements of xd_bytes bytes (int64
                                  unless opts.xd_bytes asked for a narrower transport);
                                  xd_bytes = 1: [total_records] reaction indices (uint8)   */
    int64_t* njumps;           /* [n_traj] PDMPResult.njumps (src/utils.jl:79)             */
    int64_t* nrejected;        /* [n_traj] PDMPResult.nrejected (fictitous_jumps)          */
    uint8_t* status;           /* [n_traj] PDMP_ST_*                                       */
    /* ensemble statistics at sample_times: component c in [0, n_c+n_d)                    */
    double* stat_count;        /* [T][C] trajectories that reached that sample time        */
    double* stat_sum;          /* [T][C]                                                   */
    double* stat_sumsq;        /* [T][C]                                                   */
    uint64_t* hist;            /* [T][hist_n][hist_bins], out-of-range clamps to end bins  */
    /* work counters, summed over the shard                                                */
    uint64_t rk_attempts;      /* accepted + rejected step attempts of the main integrator */
    uint64_t rk_rejects;
    uint64_t aux_attempts;     /* attempts spent in sample-time / last-bit sub-integration */
    uint64_t total_jumps;      /* executed jumps                                           */
    uint64_t total_candidates; /* thresholds (CHV) or candidate events (Rejection)         */
    double kernel_ms;          /* ensemble kernel, device time                             */
    double compact_ms;         /* pool -> CSR compaction kernel(s), device time            */
    double h2d_ms, d2h_ms;
    int32_t n_sample_times, n_comp, hist_n, hist_bins;
    int32_t n_c, n_d;
    int32_t xd_bytes, reserved0;
    void* owner;               /* private                                                  */
    /* ---- ABI v8 ---- */
    double* rate;              /* [total_records] when opts.save_rate, else NULL           */
    /* Batch statistics, for confidence intervals that assume no distribution: trajectory i belongs to
     * batch (global id) mod PDMP_STAT_BATCHES.  Sums are SHIFTED by stat_pivot[c] (the initial condition
     * of component c when it is common to all trajectories, else 0) so that variances do not cancel:
     *   batch_count[b][t][c] ; batch_sum = sum (v - pivot_c) ; batch_sumsq = sum (v - pivot_c)^2
     * stat_count / stat_sum / stat_sumsq above are the totals over the batches (un-shifted).  All sums are
     * formed in a fixed order: two runs of the same configuration return bit-identical statistics.   */
    double* stat_pivot;        /* [C]                                                      */
    double* batch_count;       /* [PDMP_STAT_BATCHES][T][C]                                */
    double* batch_sum;
    double* batch_sumsq;
    /* Multi-device runs (opts.n_devices > 1): one block per device shard, each a complete
     * pdmp_result for trajectories [traj_begin, traj_begin + n_traj) with its own CSR arrays;
     * the top-level arrays offsets/time/xc/xd/rate/njumps/nrejected/status are then NULL, the
     * counters and statistics above are the sums over the blocks (summed in block order).   */
    struct pdmp_result* blocks;
    uint64_t traj_begin;       /* global index of this block's first trajectory (0 at top level) */
    int32_t n_blocks;          /* 0: single device                                         */
    int32_t device;            /* device this result / block ran on                        */
} pdmp_result;

/* ---- entry points ------------------------------------------------------------------ */

int32_t pdmp_abi_version(void);
const char* pdmp_last_error(void);
int32_t pdmp_device_count(void);          /* 0 when no GPU is usable */

int32_t pdmp_model_count(void);
int32_t pdmp_model_info_by_id(int32_t id, pdmp_model_info* out);
int32_t pdmp_model_lookup(const char* name, pdmp_model_info* out);

/* One-shot: host buffers in, host buffers out.  Replaces "call solve(problem, CHV(ode))
 * n_traj times" (reference src/chvdiffeq.jl:211-237 / src/rejectiondiffeq.jl:155-176). */
int32_t pdmp_solve_ensemble(const pdmp_problem_desc* desc, const pdmp_solve_opts* opts,
                            pdmp_result* result);
void pdmp_result_free(pdmp_result* result);

/* Resident ensemble: device buffers and pinned host staging persist across runs, so a
 * caller doing many solves pays allocation once (the reference re-uses `problem` the same
 * way: init!(problem) src/problem.jl:150-161). */
typedef struct pdmp_ensemble pdmp_ensemble;

int32_t pdmp_ensemble_create(const pdmp_problem_desc* desc, const pdmp_solve_opts* opts,
                             pdmp_ensemble** out);
/* Re-upload initial conditions from HOST memory (xc0: [n_traj][n_c] or [n_c]; may be NULL
 * to keep the resident ones). Asynchronous on the ensemble's stream. */
int32_t pdmp_ensemble_upload_ics(pdmp_ensemble* e, const double* xc0, int32_t xc0_per_traj,
                                 const int64_t* xd0, int32_t xd0_per_traj);
/* Launch the ensemble kernel(s) with inputs resident in HBM.  `stream` is a cudaStream_t
 * passed as void*; NULL means the ensemble's own stream, so pass cudaStreamLegacy
 * ((void*)1) to order the run on the legacy default stream.  Asynchronous: the run starts
 * after the work already queued on `stream`, and work queued on it later sees the results
 * (that is how the caller's NCCL all-reduce of the statistics is ordered). */
int32_t pdmp_ensemble_run(pdmp_ensemble* e, uint64_t seed, void* stream);
/* The host-buffer step: upload initial conditions from HOST memory (NULL: keep the resident
 * ones), run, and stream every chunk's results to pinned host staging while later chunks still
 * compute (D2H overlaps the kernels).  Synchronous; `result` holds HOST views valid until the
 * next run/destroy.  This is what pdmp_solve_ensemble does after creating the ensemble. */
int32_t pdmp_ensemble_run_host(pdmp_ensemble* e, uint64_t seed, const double* xc0, int32_t xc0_per_traj,
                               const int64_t* xd0, int32_t xd0_per_traj, int32_t fetch_records,
                               pdmp_result* result);
/* Wait for the run and fill `result` with HOST views (pinned staging owned by the
 * ensemble, valid until the next run/destroy).  fetch_records = 0 skips the trajectory
 * arrays (counts, status, statistics and counters only). */
int32_t pdmp_ensemble_fetch(pdmp_ensemble* e, int32_t fetch_records, pdmp_result* result);
/* Device pointers of the statistics buffers, for the caller's NCCL all-reduce: the batch buffer
 * [3][PDMP_STAT_BATCHES][T][C] doubles (count | shifted sum | shifted sum of squares: additive across
 * shards, the pivot being a property of the problem) and hist [T][hist_n][hist_bins] uint64.  A fetch after
 * the all-reduce returns the global statistics. */
int32_t pdmp_ensemble_stats_device(pdmp_ensemble* e, void** stats_f64, uint64_t* n_f64,
                                   void** hist_u64, uint64_t* n_u64);
/* Work counters of the last run without fetching anything else (synchronises). */
int32_t pdmp_ensemble_counters(pdmp_ensemble* e, pdmp_result* result);
void pdmp_ensemble_destroy(pdmp_ensemble* e);

/* Measured FP64 FMA peak of `device` in TFLOP/s (dependent-chain-free DFMA kernel), the
 * roofline denominator MEASURED_PEAKS.json does not carry. */
int32_t pdmp_fp64_peak_tflops(int32_t device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* PDMP_B200_H */
