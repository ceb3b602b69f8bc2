// stats_kernel.cuh — ensemble moments and histograms at the fixed sample times, from the rows the ensemble
// kernel recorded (ensemble_kernel.cuh, Traj::accumulate): one row of C doubles per (trajectory, sample time).
//
// The reference has no ensemble API; this is the north star's "on-GPU ensemble moment and histogram
// accumulators".  Design goals:
//   * reproducible: every sum is formed in a FIXED order (per-thread strides, shuffle butterflies, warp order),
//     so two runs of the same configuration return bit-identical statistics — no floating-point atomics;
//   * confidence intervals without a distributional assumption: trajectories are split into
//     PDMP_STAT_BATCHES = 64 batches by (global id) mod 64 and every batch keeps its own count / sum / sum of
//     squares, so that the host can form batch means (CI of the mean) and batch variances (CI of the variance);
//   * no cancellation: sums are SHIFTED by a pivot per component (the problem's initial condition), the
//     variance comes out of sum (v - p)^2 and sum (v - p);
//   * additive across chunks, device shards and ranks (same pivot everywhere): the multi-GPU reduction stays a
//     plain sum (NCCL all-reduce of the batch buffer, or the in-library block sum).
// HBM-bound: it reads every recorded row once (TCP at the bench size: 10^7 x 4 x 3 doubles = 0.96 GB, ~0.3 ms).
#pragma once
#include "ensemble_kernel.cuh"

namespace pdmp {

constexpr int STAT_BATCHES = PDMP_STAT_BATCHES;
constexpr int STAT_MAXC = 2 * MAXD;   // C = NC + ND components per row

struct StatParams {
    const double* samples;          // [n][T][C] rows of this launch
    unsigned long long n;           // trajectories of this launch
    unsigned long long id0;         // global id of local trajectory 0 (batch = global id mod 64)
    int T, C, H, bins;
    int hist_comp[MAXH];
    double hist_lo[MAXH], hist_scale[MAXH];   // bin = floor((v - lo) * scale), clamped to the end bins
    double pivot[STAT_MAXC];
    double* out;                    // [3][64][T][C]: count | sum (v - p) | sum (v - p)^2 ; += (chunks add up)
    unsigned long long* hist;       // [T][H][bins] ; +=
};

// grid = (64 batches, T sample times), 256 threads.  Thread t of CTA (b, j) visits the trajectories of batch b
// in the fixed order t, t + 256, t + 512, ... (counted inside the batch), so its partial sums do not depend on
// scheduling; the CTA reduction is a shuffle butterfly per warp followed by the warps in order.
__global__ void __launch_bounds__(256) stats_reduce_kernel(const StatParams S) {
    __shared__ double s_part[8][1 + 2 * STAT_MAXC];
    extern __shared__ unsigned s_hist[];   // [H][bins]
    const int b = blockIdx.x, j = blockIdx.y, C = S.C;
    const int nh = S.H * S.bins;
    for (int i = threadIdx.x; i < nh; i += 256) s_hist[i] = 0u;
    if (nh) __syncthreads();
    double cnt = 0.0, s1[STAT_MAXC], s2[STAT_MAXC];
#pragma unroll
    for (int c = 0; c < STAT_MAXC; ++c) { s1[c] = 0.0; s2[c] = 0.0; }
    const unsigned long long first = (unsigned long long)((b + STAT_BATCHES - (int)(S.id0 % STAT_BATCHES)) % STAT_BATCHES);
    for (unsigned long long i = first + (unsigned long long)STAT_BATCHES * threadIdx.x; i < S.n; i += (unsigned long long)STAT_BATCHES * 256) {
        const double* row = S.samples + (i * S.T + j) * C;
        const double v0 = __ldcs(row);
        if (v0 == v0) {   // NaN in slot 0: the trajectory never reached this sample time
            cnt += 1.0;
            double v[STAT_MAXC];
#pragma unroll
            for (int c = 0; c < STAT_MAXC; ++c)
                if (c < C) {
                    v[c] = c ? __ldcs(row + c) : v0;
                    const double d = v[c] - S.pivot[c];
                    s1[c] += d;
                    s2[c] = fma(d, d, s2[c]);
                }
            for (int h = 0; h < S.H; ++h) {
                double x = 0.0;
#pragma unroll
                for (int c = 0; c < STAT_MAXC; ++c) if (c == S.hist_comp[h]) x = v[c];
                const double fb = floor((x - S.hist_lo[h]) * S.hist_scale[h]);
                int bin = fb < 0.0 ? 0 : (fb >= (double)S.bins ? S.bins - 1 : (int)fb);
                if (!(fb == fb)) bin = 0;
                atomicAdd(&s_hist[h * S.bins + bin], 1u);   // integers: order-independent
            }
        }
    }
    // CTA reduction in a fixed shape
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    auto wsum = [](double x) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        return x;
    };
    cnt = wsum(cnt);
    if (lane == 0) s_part[w][0] = cnt;
#pragma unroll
    for (int c = 0; c < STAT_MAXC; ++c)
        if (c < C) {
            const double a = wsum(s1[c]), q = wsum(s2[c]);
            if (lane == 0) { s_part[w][1 + c] = a; s_part[w][1 + STAT_MAXC + c] = q; }
        }
    __syncthreads();
    const size_t TC = (size_t)S.T * C, plane = (size_t)STAT_BATCHES * TC;
    if (threadIdx.x < C) {
        const int c = threadIdx.x;
        double n = 0.0, a = 0.0, q = 0.0;
        for (int k = 0; k < 8; ++k) { n += s_part[k][0]; a += s_part[k][1 + c]; q += s_part[k][1 + STAT_MAXC + c]; }
        const size_t at = (size_t)b * TC + (size_t)j * C + c;
        S.out[at] += n;                 // one writer per element and launch; launches are ordered by the caller
        S.out[plane + at] += a;
        S.out[2 * plane + at] += q;
    }
    for (int i = threadIdx.x; i < nh; i += 256)
        if (s_hist[i]) atomicAdd(S.hist + (size_t)j * nh + i, (unsigned long long)s_hist[i]);
}

}  // namespace pdmp
