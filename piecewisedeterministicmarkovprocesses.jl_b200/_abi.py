"""ctypes mirror of include/pdmp_b200.h (ABI version 8).

Pure declarations: no library is loaded here, so the oracle loader (oracle/oracle.py, test
infrastructure) can share the struct layouts without touching the product library.
"""
import ctypes as C

ABI_VERSION = 8

ALG_CHV, ALG_REJECTION, ALG_REJECTION_EXACT = 0, 1, 2
ODE_DP5, ODE_TSIT5 = 0, 1
STAT_BATCHES = 64
RNG_PHILOX, RNG_REPLAY = 0, 1
PREC_FP64, PREC_FP32 = 0, 1

ST_OK, ST_HIT_NJUMPS, ST_NO_PROGRESS, ST_NONPOSITIVE_RATE, ST_BOUND_VIOLATED, ST_NAN, \
    ST_MAX_ATTEMPTS, ST_OUTPUT_OVERFLOW, ST_REPLAY_EXHAUSTED = range(9)
STATUS_NAMES = ["OK", "HIT_NJUMPS", "NO_PROGRESS", "NONPOSITIVE_RATE", "BOUND_VIOLATED", "NAN",
                "MAX_ATTEMPTS", "OUTPUT_OVERFLOW", "REPLAY_EXHAUSTED"]

OK, ERR_INVALID, ERR_MODEL, ERR_CUDA, ERR_NO_DEVICE, ERR_CAPACITY, ERR_UNSUPPORTED = range(7)

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_uint64_p = C.POINTER(C.c_uint64)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)


class ModelInfo(C.Structure):
    _fields_ = [("id", C.c_int32), ("n_c", C.c_int32), ("n_d", C.c_int32), ("n_r", C.c_int32),
                ("n_parms", C.c_int32), ("has_bound", C.c_int32), ("has_delta", C.c_int32),
                ("zero_flow", C.c_int32), ("exact_flow", C.c_int32), ("composite", C.c_int32),
                ("name", C.c_char * 32)]


class ProblemDesc(C.Structure):
    _fields_ = [("n_traj", C.c_uint64), ("traj_id_offset", C.c_uint64),
                ("xc0", c_double_p), ("xd0", c_int64_p), ("nu", c_int64_p), ("parms", c_double_p),
                ("t0", C.c_double), ("tf", C.c_double),
                ("model_id", C.c_int32), ("xc0_per_traj", C.c_int32), ("xd0_per_traj", C.c_int32),
                ("nu_col_major", C.c_int32), ("n_parms", C.c_int32), ("rate_kind", C.c_int32)]


class SolveOpts(C.Structure):
    _fields_ = [("reltol", C.c_double), ("abstol", C.c_double),
                ("n_jumps", C.c_int64), ("max_attempts", C.c_int64), ("seed", C.c_uint64),
                ("replay_u", c_double_p), ("replay_stride", C.c_uint64),
                ("sample_times", c_double_p), ("hist_comp", c_int32_p),
                ("hist_lo", c_double_p), ("hist_hi", c_double_p), ("max_records", C.c_uint64),
                ("algorithm", C.c_int32), ("ode", C.c_int32), ("precision", C.c_int32),
                ("rng_mode", C.c_int32), ("save_pre", C.c_int32), ("save_post", C.c_int32),
                ("no_records", C.c_int32), ("n_sample_times", C.c_int32), ("hist_n", C.c_int32),
                ("hist_bins", C.c_int32), ("device", C.c_int32), ("tstop_policy", C.c_int32),
                ("ref_quirks", C.c_int32), ("xd_bytes", C.c_int32),
                ("save_c_count", C.c_int32), ("save_d_count", C.c_int32),
                ("devices", c_int32_p), ("save_c_mask", C.c_uint64), ("save_d_mask", C.c_uint64),
                ("n_devices", C.c_int32), ("save_rate", C.c_int32)]


class Result(C.Structure):
    pass


Result._fields_ = [("n_traj", C.c_uint64), ("total_records", C.c_uint64), ("records_needed", C.c_uint64),
                ("offsets", c_uint64_p), ("time", c_double_p), ("xc", c_double_p), ("xd", C.c_void_p),
                ("njumps", c_int64_p), ("nrejected", c_int64_p), ("status", c_uint8_p),
                ("stat_count", c_double_p), ("stat_sum", c_double_p), ("stat_sumsq", c_double_p),
                ("hist", c_uint64_p),
                ("rk_attempts", C.c_uint64), ("rk_rejects", C.c_uint64), ("aux_attempts", C.c_uint64),
                ("total_jumps", C.c_uint64), ("total_candidates", C.c_uint64),
                ("kernel_ms", C.c_double), ("compact_ms", C.c_double),
                ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("n_sample_times", C.c_int32), ("n_comp", C.c_int32), ("hist_n", C.c_int32),
                ("hist_bins", C.c_int32), ("n_c", C.c_int32), ("n_d", C.c_int32),
                ("xd_bytes", C.c_int32), ("reserved0", C.c_int32),
                ("owner", C.c_void_p),
                ("rate", c_double_p),
                ("stat_pivot", c_double_p), ("batch_count", c_double_p), ("batch_sum", c_double_p), ("batch_sumsq", c_double_p),
                ("blocks", C.POINTER(Result)), ("traj_begin", C.c_uint64),
                ("n_blocks", C.c_int32), ("device", C.c_int32)]


# every symbol include/pdmp_b200.h declares (tests check the built library exports all)
EXPORTS = [
    "pdmp_abi_version", "pdmp_last_error", "pdmp_device_count", "pdmp_model_count",
    "pdmp_model_info_by_id", "pdmp_model_lookup", "pdmp_solve_ensemble", "pdmp_result_free",
    "pdmp_ensemble_create", "pdmp_ensemble_upload_ics", "pdmp_ensemble_run", "pdmp_ensemble_run_host", "pdmp_ensemble_fetch",
    "pdmp_ensemble_stats_device", "pdmp_ensemble_counters", "pdmp_ensemble_destroy",
    "pdmp_fp64_peak_tflops",
]
