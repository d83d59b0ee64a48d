// epilogue.cuh -- SURVEY.md 8(f) rank 4: what Setup_results and PairWiseComparisons do AFTER the per-gene statistics,
// on the device, for the 1M-row scale (C5) where the host versions are the wall clock:
//   * the stable sort of the tested genes by naive p (scoary/methods.py:903, :1448-1454: sorted(), ties keep
//     dict-insertion order = gene order),
//   * Bonferroni  min(p m, 1)                                        (methods.py:923-924),
//   * Benjamini-Hochberg step-up with the reference's tie rule: the least significant gene keeps its p, a gene
//     tied with its less significant neighbour inherits that neighbour's value, everything else takes
//     min(p m / rank, value of the next less significant gene)      (methods.py:903-919),
//   * the two-sided exact binomial test at p = 0.5 of the pair counts (ss.binom_test call sites, methods.py:1267-1275).
// The arithmetic is the reference's, operation for operation (one correctly rounded multiplication, then one division),
// so Bonferroni / BH columns equal the host implementation bit for bit; the binomial p is a direct sum over the
// log-factorial LUT (<= 1e-13 relative to SciPy's incomplete-beta value).
//
// Kernels: an LSD radix sort written for this job (64-bit keys = the bit pattern of p, 8-bit digits, one warp per
// 1 024 consecutive keys; a warp ranks its keys in order with __match_any_sync, so every pass is stable), a two-level
// suffix-min scan, and two elementwise kernels.  HBM-bound, tiny next to the walks; no library calls.
#pragma once
#include "common.cuh"

namespace sb {

constexpr int RADIX_BITS = 8, RADIX = 1 << RADIX_BITS;
constexpr int SORT_CHUNK = 1024;          // keys per warp
constexpr int SORT_WARPS = 8;             // warps per block

// keys[i] = bits of p[i] for tested genes (p >= 0: the IEEE pattern orders like the value), ~0 for the others so that
// they sort behind every tested gene; idx[i] = i.  *n_kept counts the tested genes.
__global__ void __launch_bounds__(256) epi_keys_kernel(const double *__restrict__ p, const int32_t *__restrict__ counts,
                                                       const uint8_t *__restrict__ keep, int64_t n,
                                                       uint64_t *__restrict__ keys, int32_t *__restrict__ idx,
                                                       unsigned long long *__restrict__ n_kept)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool k = false;
    if (i < n) {
        // the skip rule of Setup_results (methods.py:804-814): genes present in all or in none of the isolates that
        // have a value for the trait are not tested
        k = keep ? keep[i] != 0
                 : (counts[i * 4 + 0] + counts[i * 4 + 1] > 0) && (counts[i * 4 + 2] + counts[i * 4 + 3] > 0);
        keys[i] = k ? (uint64_t)__double_as_longlong(p[i]) : ~0ULL;
        idx[i] = (int32_t)i;
    }
    const unsigned m = __ballot_sync(0xffffffffu, k);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_kept, (unsigned long long)__popc(m));
}

// hist[d * n_chunks + c] = keys of chunk c whose digit (key >> shift) & 255 is d
__global__ void __launch_bounds__(32 * SORT_WARPS) radix_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift,
                                                                    int n_chunks, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t s_h[SORT_WARPS][RADIX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.x * SORT_WARPS + warp;
    for (int d = lane; d < RADIX; d += 32) s_h[warp][d] = 0;
    __syncwarp();
    if (chunk < n_chunks) {
        const int64_t lo = (int64_t)chunk * SORT_CHUNK, hi = min(n, lo + SORT_CHUNK);
        for (int64_t i = lo + lane; i < hi; i += 32) atomicAdd(&s_h[warp][(keys[i] >> shift) & (RADIX - 1)], 1u);
        __syncwarp();
        for (int d = lane; d < RADIX; d += 32) hist[(int64_t)d * n_chunks + chunk] = s_h[warp][d];
    }
}

// exclusive scan of hist in (digit, chunk) order, in place; one block
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t *__restrict__ hist, int64_t total)
{
    __shared__ uint32_t s_part[1024];
    const int tid = threadIdx.x;
    const int64_t per = (total + 1023) / 1024;
    const int64_t lo = min(total, (int64_t)tid * per), hi = min(total, lo + per);
    uint32_t sum = 0;
    for (int64_t i = lo; i < hi; ++i) sum += hist[i];
    s_part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {      // Hillis-Steele inclusive scan of the partial sums
        const uint32_t v = tid >= o ? s_part[tid - o] : 0u;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    uint32_t run = s_part[tid] - sum;
    for (int64_t i = lo; i < hi; ++i) {
        const uint32_t v = hist[i];
        hist[i] = run;
        run += v;
    }
}

// stable scatter: a warp walks its chunk in order, 32 keys at a time
__global__ void __launch_bounds__(32 * SORT_WARPS) radix_scatter_kernel(const uint64_t *__restrict__ keys_in,
                                                                       const int32_t *__restrict__ idx_in, int64_t n,
                                                                       int shift, int n_chunks,
                                                                       const uint32_t *__restrict__ offs,
                                                                       uint64_t *__restrict__ keys_out,
                                                                       int32_t *__restrict__ idx_out)
{
    __shared__ uint32_t s_o[SORT_WARPS][RADIX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.x * SORT_WARPS + warp;
    if (chunk >= n_chunks) return;
    for (int d = lane; d < RADIX; d += 32) s_o[warp][d] = offs[(int64_t)d * n_chunks + chunk];
    __syncwarp();
    const int64_t lo = (int64_t)chunk * SORT_CHUNK, hi = min(n, lo + SORT_CHUNK);
    for (int64_t base = lo; base < hi; base += 32) {
        const int64_t i = base + lane;
        const bool live = i < hi;
        const uint64_t key = live ? keys_in[i] : 0;
        const int d = live ? (int)((key >> shift) & (RADIX - 1)) : -1 - lane;      // idle lanes match nobody
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t dst = 0;
        if (live) dst = s_o[warp][d] + rank;
        __syncwarp();
        if (live && rank == 0) s_o[warp][d] += __popc(peers);
        __syncwarp();
        if (live) {
            keys_out[dst] = key;
            idx_out[dst] = idx_in[i];
        }
    }
}

// vals[i] for the suffix-min scan (methods.py:903-919), i = rank - 1 in the sorted order:
//   the last one keeps its p; tied with the next -> +inf (it inherits whatever the next one gets); else p m / rank
__global__ void __launch_bounds__(256) bh_values_kernel(const uint64_t *__restrict__ keys_sorted, int64_t m,
                                                        double n_tests, double *__restrict__ vals)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double p = __longlong_as_double((long long)keys_sorted[i]);
    double v;
    if (i == m - 1) v = p;
    else if (keys_sorted[i] == keys_sorted[i + 1]) v = __longlong_as_double(0x7FF0000000000000LL);
    else v = __ddiv_rn(__dmul_rn(p, n_tests), (double)(i + 1));
    vals[i] = v;
}

// suffix minimum, level 1: within blocks of 1024; block_min[b] = min of block b
__global__ void __launch_bounds__(1024) suffix_min_block_kernel(double *__restrict__ vals, int64_t m,
                                                                double *__restrict__ block_min)
{
    __shared__ double s[1024];
    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * 1024 + tid;
    s[tid] = i < m ? vals[i] : __longlong_as_double(0x7FF0000000000000LL);
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const double v = tid + o < 1024 ? s[tid + o] : __longlong_as_double(0x7FF0000000000000LL);
        __syncthreads();
        s[tid] = fmin(s[tid], v);
        __syncthreads();
    }
    if (i < m) vals[i] = s[tid];
    if (tid == 0) block_min[blockIdx.x] = s[0];
}

// level 2: block_min[b] <- min over blocks > b (exclusive suffix min); one block
__global__ void __launch_bounds__(1024) suffix_min_top_kernel(double *__restrict__ block_min, int64_t n_blocks)
{
    __shared__ double s_part[1024];
    const int tid = threadIdx.x;
    const double INF = __longlong_as_double(0x7FF0000000000000LL);
    const int64_t per = (n_blocks + 1023) / 1024;
    const int64_t lo = min(n_blocks, (int64_t)tid * per), hi = min(n_blocks, lo + per);
    double mn = INF;
    for (int64_t i = lo; i < hi; ++i) mn = fmin(mn, block_min[i]);
    s_part[tid] = mn;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {      // inclusive suffix min of the partials
        const double v = tid + o < 1024 ? s_part[tid + o] : INF;
        __syncthreads();
        s_part[tid] = fmin(s_part[tid], v);
        __syncthreads();
    }
    double run = tid + 1 < 1024 ? s_part[tid + 1] : INF;      // everything right of this thread's range
    for (int64_t i = hi - 1; i >= lo; --i) {
        const double v = block_min[i];
        block_min[i] = run;
        run = fmin(run, v);
    }
}

// bh[gene] = min(min(suffix-min inside the block, blocks to the right), 1); bonferroni[gene] = min(p m, 1);
// order[i] = gene at rank i.  Untested genes get NaN in both columns.
__global__ void __launch_bounds__(256) bh_finish_kernel(const uint64_t *__restrict__ keys_sorted,
                                                        const int32_t *__restrict__ idx_sorted, int64_t n, int64_t m,
                                                        double n_tests, const double *__restrict__ vals,
                                                        const double *__restrict__ block_min,
                                                        int32_t *__restrict__ order, double *__restrict__ bonferroni,
                                                        double *__restrict__ bh)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t g = idx_sorted[i];
    if (order) order[i] = g;
    if (i < m) {
        const double p = __longlong_as_double((long long)keys_sorted[i]);
        if (bh) bh[g] = fmin(fmin(vals[i], block_min[i >> 10]), 1.0);
        if (bonferroni) bonferroni[g] = fmin(__dmul_rn(p, n_tests), 1.0);
    } else {
        const double nan = __longlong_as_double(0x7FF8000000000000LL);
        if (bh) bh[g] = nan;
        if (bonferroni) bonferroni[g] = nan;
    }
}

// ss.binom_test(k, n, 0.5) (methods.py:1267-1275; SciPy's binomtest, two-sided): with p = 0.5 the distribution is
// symmetric, so the terms no more likely than pmf(k) are the two tails from min(k, n - k) outwards:
//   p = 1 if 2 k == n, else min(1, 2 sum_{i <= min(k, n-k)} C(n, i) / 2^n);   n = 0 has no answer (NaN, as the host path).
// The sum runs from the largest term down until the terms fall below 2^-60 of it; the largest term comes from the
// double-double LUT (one exp), the others from the ratio C(n, i-1) / C(n, i) = i / (n - i + 1).
__global__ void __launch_bounds__(256) binom_two_sided_kernel(const int32_t *__restrict__ k_arr,
                                                              const int32_t *__restrict__ n_arr, int64_t stride,
                                                              int64_t count, const double2 *__restrict__ lut,
                                                              double *__restrict__ out)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    const int n = n_arr[e * stride];
    int k = k_arr[e * stride];
    if (n <= 0 || k < 0 || k > n) { out[e] = __longlong_as_double(0x7FF8000000000000LL); return; }
    k = min(k, n - k);
    if (2 * k == n) { out[e] = 1.0; return; }
    dd L = dd_sub(dd_make(lut[n]), dd_add(dd_make(lut[k]), dd_make(lut[n - k])));
    const double ln2_hi = 0.6931471805599453, ln2_lo = 2.3190468138462996e-17;
    const double nl = (double)n * ln2_hi;                          // n ln 2 as an exact product (hi + rounding error) + low part
    L = dd_sub(L, dd{nl, fma((double)n, ln2_hi, -nl) + (double)n * ln2_lo});
    const double top = exp(L.hi) * (1.0 + L.lo);
    double term = top, sum = top;
    for (int i = k; i >= 1; --i) {
        term = term * ((double)i / (double)(n - i + 1));
        sum += term;
        if (term < top * 8.673617379884035e-19) break;
    }
    out[e] = fmin(1.0, 2.0 * sum);
}

}  // namespace sb
