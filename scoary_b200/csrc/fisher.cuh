// fisher.cuh -- K2+K3: per-gene 2x2 contingency table by popcount over the
// packed bitset rows, then the two-sided Fisher exact p from a double-double
// log-factorial LUT.
//
// Replaces, per (gene, trait): Perform_statistics (scoary/methods.py:930-982)
// and the ss.fisher_exact call (methods.py:842-857).  The p-value rule is
// SciPy 1.18.1's (scipy/stats/_stats_py.py, fisher_exact, two-sided):
//   p = sum_{x on the observed side, from the extreme up to a} pmf(x)
//     + sum_{x on the other side of the mode with pmf(x) <= pmf(a)(1+1e-14)} pmf(x)
//   p = 1 when pmf(a) ~= pmf(mode);  p = min(p, 1).
//
// Data movement: a persistent grid; every warp streams its own gene rows into
// shared memory with 1-D TMA bulk copies (cp.async.bulk + its own pair of
// mbarriers, double buffered) and owns one row at a time: 128-bit shared
// loads, __popcll, warp REDUX, then the warp walks the hypergeometric support.
#pragma once
#include "common.cuh"

// Warp collectives and rounding intrinsics of the Fisher code: CUDA spellings in a device build, the emulation
// driver's versions (32 host threads in lockstep) under SB_HOST_EMUL.
#ifdef SB_HOST_EMUL
#define SB_SHFL_XOR(v, o) sb_emul_shfl_xor(v, o)
#define SB_SHFL(v, src) sb_emul_shfl(v, src)
#define SB_BALLOT(pred) sb_emul_ballot(pred)
#define SB_FFS(x) __builtin_ffs((int)(x))
#define SB_DMUL(a, b) sb_emul_mul(a, b)
#define SB_DDIV(a, b) sb_emul_div(a, b)
#else
#define SB_SHFL_XOR(v, o) __shfl_xor_sync(0xffffffffu, v, o)
#define SB_SHFL(v, src) __shfl_sync(0xffffffffu, v, src)
#define SB_BALLOT(pred) __ballot_sync(0xffffffffu, pred)
#define SB_FFS(x) __ffs(x)
#define SB_DMUL(a, b) __dmul_rn(a, b)
#define SB_DDIV(a, b) __ddiv_rn(a, b)
#endif

namespace sb {

#ifndef SB_FISHER_THREADS
#define SB_FISHER_THREADS 1024
#endif
constexpr int FISHER_THREADS = SB_FISHER_THREADS;
constexpr double FISHER_TIE_TOL = 1e-12;   // |log pmf ratio| below this is a tie (exact ties give 0)

struct FisherArgs {
    const uint64_t *genes;   // [G][W]
    int64_t G;
    int32_t W;               // words per row (even)
    int32_t Wn;              // ceil(N/64): words that carry isolates (hash domain)
    const uint64_t *tvalue;  // [W]
    const uint64_t *tmask;   // [W]
    const double2 *lut;      // [lut_n + 1] log k! as (hi, lo)
    int32_t lut_n;
    int32_t *counts;         // [G][4] or null
    double *p;               // [G] or null
    uint64_t *hash;          // [G][2] or null
};

// S(x) = lf[x] + lf[n1-x] + lf[n-x] + lf[n2-n+x]  (the x-dependent part of -log pmf)
SB_DEV dd fisher_S(const double2 *lut, int x, int n1, int n2, int n)
{
    dd s = dd_make(lut[x]);
    s = dd_add(s, dd_make(lut[n1 - x]));
    s = dd_add(s, dd_make(lut[n - x]));
    s = dd_add(s, dd_make(lut[n2 - n + x]));
    return s;
}

SB_DEV double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += SB_SHFL_XOR(v, o);
    return v;
}

// Sum pmf(x) for x = x0, x0+dir, ..., count terms, terms non-increasing.
// logp_a = log pmf(a) (dd), S_a = S(a).  Each lane takes FISHER_BLOCK consecutive terms per round:
// the first from the double-double LUT (one exp), the rest by the hypergeometric ratio
//   pmf(x+1)/pmf(x) = (n1-x)(n-x) / ((x+1)(n2-n+x+1))       (and its mirror for dir = -1)
// in plain double (two roundings per step: <= 2e-15 relative after 7 steps).  Stops once a round
// starts below 2^-80 pmf(a).
constexpr int FISHER_BLOCK = 8;

SB_DEV double fisher_tail(const double2 *lut, int x0, int dir, int count, int n1, int n2,
                                              int n, dd logp_a, dd S_a, double pexact, int lane)
{
    double acc = 0.0;
    const double cut = pexact * 8.271806125530277e-25;   // 2^-80
    for (int base = 0; base < count; base += 32 * FISHER_BLOCK) {
        const int kb = base + lane * FISHER_BLOCK;
        double t = 0.0;
        if (kb < count) {
            const int x = x0 + dir * kb;
            dd d = dd_sub(S_a, fisher_S(lut, x, n1, n2, n));   // log pmf(x) - log pmf(a)
            dd L = dd_add(logp_a, d);
            t = exp(L.hi) * (1.0 + L.lo);
            double A, B, C, D;      // pmf(next)/pmf(cur) = (A * B) / (C * D); A, B step down, C, D step up
            if (dir > 0) { A = (double)(n1 - x); B = (double)(n - x); C = (double)(x + 1); D = (double)(n2 - n + x + 1); }
            else { A = (double)x; B = (double)(n2 - n + x); C = (double)(n1 - x + 1); D = (double)(n - x + 1); }
            double tt = t, s = t;
#pragma unroll
            for (int j = 1; j < FISHER_BLOCK; ++j) {
                if (kb + j < count) {
                    tt = SB_DMUL(tt, SB_DDIV(SB_DMUL(A, B), SB_DMUL(C, D)));
                    s += tt;
                    A -= 1.0; B -= 1.0; C += 1.0; D += 1.0;
                }
            }
            acc += s;
        }
        const double first = SB_SHFL(t, 0);
        if (first < cut) break;
    }
    return warp_sum(acc);
}

// Warp-cooperative two-sided Fisher exact p for [[a, b], [c, d]].
SB_DEV double fisher_two_sided_warp(const double2 *lut, int a, int b, int c, int d, int lane)
{
    if (a + b == 0 || c + d == 0 || a + c == 0 || b + d == 0) return 1.0;
    // The p-value is invariant under swapping rows, swapping columns and transposing.  Put
    // the table in a canonical orientation first so that all eight variants run the very
    // same arithmetic and return bit-identical p (ties stay ties for the BH tie rule and
    // the p-sort).  Invariant: the unordered pair of unordered pairs {{a,d},{b,c}}.
    {
        int d0 = min(a, d), d1 = max(a, d), o0 = min(b, c), o1 = max(b, c);
        if (o0 < d0 || (o0 == d0 && o1 < d1)) {   // the lexicographically smaller pair goes on the diagonal
            int t0 = d0, t1 = d1;
            d0 = o0; d1 = o1; o0 = t0; o1 = t1;
        }
        a = d0; d = d1; b = o0; c = o1;
    }
    const int n1 = a + b, n2 = c + d, n = a + c, M = n1 + n2;
    const int lo = max(0, n - n2), hi = min(n, n1);
    const int mode = (int)((double)((long long)(n + 1) * (long long)(n1 + 1)) / (double)(M + 2));
    if (a == mode) return 1.0;
    // log pmf(a)
    dd base = dd_make(lut[n1]);
    base = dd_add(base, dd_make(lut[n2]));
    base = dd_add(base, dd_make(lut[n]));
    base = dd_add(base, dd_make(lut[M - n]));
    base = dd_sub(base, dd_make(lut[M]));
    const dd S_a = fisher_S(lut, a, n1, n2, n);
    const dd logp_a = dd_sub(base, S_a);
    const double pexact = exp(logp_a.hi) * (1.0 + logp_a.lo);
    {   // pexact ~= pmode  ->  1
        dd dm = dd_sub(S_a, fisher_S(lut, mode, n1, n2, n));
        if (fabs(dm.hi + dm.lo) <= FISHER_TIE_TOL) return 1.0;
    }
    const int dir_obs = (a < mode) ? -1 : +1;          // away from the mode on the observed side
    const int cnt_obs = (a < mode) ? (a - lo + 1) : (hi - a + 1);
    double p = fisher_tail(lut, a, dir_obs, cnt_obs, n1, n2, n, logp_a, S_a, pexact, lane);

    // other side: y_k = mode + dir2 * k, k = 1..K; included iff log pmf(y_k) - log pmf(a) <= tol.
    // pmf decreases with k, so find the first included k with a 32-ary search.
    const int dir2 = -dir_obs;
    const int K = (dir2 > 0) ? (hi - mode) : (mode - lo);
    int lo_k = 1, hi_k = K + 1;
    while (hi_k > lo_k) {
        const int len = hi_k - lo_k;
        const int stride = (len + 31) >> 5;
        const int k = lo_k + lane * stride;
        bool pred = false;
        if (k < hi_k) {
            dd dk = dd_sub(S_a, fisher_S(lut, mode + dir2 * k, n1, n2, n));
            pred = (dk.hi + dk.lo) <= FISHER_TIE_TOL;
        }
        const unsigned ball = SB_BALLOT(pred);
        if (ball == 0u) {
            lo_k = lo_k + ((len - 1) / stride) * stride + 1;
        } else {
            const int f = SB_FFS(ball) - 1;
            hi_k = lo_k + f * stride;
            lo_k = (f == 0) ? hi_k : (lo_k + (f - 1) * stride + 1);
        }
    }
    const int kstar = lo_k;
    if (kstar <= K)
        p += fisher_tail(lut, mode + dir2 * kstar, dir2, K - kstar + 1, n1, n2, n, logp_a, S_a, pexact, lane);
    return fmin(p, 1.0);
}

#ifndef SB_HOST_EMUL   // the kernel itself (TMA pipeline, popcounts) is device-only
template <bool LUT_SMEM, bool HASH>
__global__ void __launch_bounds__(FISHER_THREADS) fisher_kernel(const FisherArgs A)
{
    // Shared memory: [2 mbarriers per warp] [value & mask] [mask] [2 row buffers per warp] [LUT].
    // Every warp runs its own double-buffered TMA pipeline over its own rows (1-D bulk copies of
    // one 8*W-byte row, completion on the warp's mbarriers): the rows cost very different amounts
    // of Fisher work, so there is no block-wide barrier anywhere in the loop.
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NW = FISHER_THREADS / 32;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);                    // [NW][2]
    uint64_t *s_tm = reinterpret_cast<uint64_t *>(smem_raw + 16 * NW);          // value & mask  [W]
    uint64_t *s_m = s_tm + A.W;                                                 // mask          [W]
    uint64_t *s_rows = s_m + A.W;                                               // [NW][2][W]
    double2 *s_lut = reinterpret_cast<double2 *>(s_rows + (size_t)NW * 2 * A.W);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t *my_bar = bars + warp * 2;
    uint64_t *my_rows = s_rows + (size_t)warp * 2 * A.W;
    const uint32_t row_bytes = (uint32_t)A.W * 8u;
    const int64_t stride = (int64_t)gridDim.x * NW;
    const int64_t row_first = (int64_t)blockIdx.x * NW + warp;

    if (lane == 0) {
        mbar_init(&my_bar[0], 1);
        mbar_init(&my_bar[1], 1);
        fence_barrier_init();
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int64_t r = row_first + b * stride;
            if (r < A.G) {
                mbar_arrive_expect_tx(&my_bar[b], row_bytes);
                tma_bulk_g2s(my_rows + (size_t)b * A.W, A.genes + r * A.W, row_bytes, &my_bar[b]);
            }
        }
    }
    for (int w = tid; w < A.W; w += FISHER_THREADS) {
        uint64_t m = A.tmask[w];
        s_m[w] = m;
        s_tm[w] = A.tvalue[w] & m;
    }
    if (LUT_SMEM) {
        for (int k = tid; k <= A.lut_n; k += FISHER_THREADS) s_lut[k] = A.lut[k];
    }
    __syncthreads();   // trait vectors and LUT staged (the only block-wide barrier)
    const double2 *lut = LUT_SMEM ? s_lut : A.lut;

    // trait totals (every warp computes them redundantly: W is tiny)
    int n_tp = 0, n_m = 0;
    for (int w = lane; w < A.W; w += 32) {
        n_tp += __popcll(s_tm[w]);
        n_m += __popcll(s_m[w]);
    }
    n_tp = __reduce_add_sync(0xffffffffu, n_tp);
    n_m = __reduce_add_sync(0xffffffffu, n_m);

    const int W2 = A.W >> 1;
    const ulonglong2 *tm2 = reinterpret_cast<const ulonglong2 *>(s_tm);
    const ulonglong2 *m2 = reinterpret_cast<const ulonglong2 *>(s_m);

    int it = 0;
    for (int64_t g_idx = row_first; g_idx < A.G; g_idx += stride, ++it) {
        const int buf = it & 1;
        mbar_wait(&my_bar[buf], (uint32_t)((it >> 1) & 1));
        const ulonglong2 *row2 = reinterpret_cast<const ulonglong2 *>(my_rows + (size_t)buf * A.W);
        int tp = 0, gp = 0;
        uint64_t h0 = 0, h1 = 0;
        for (int cidx = lane; cidx < W2; cidx += 32) {
            const ulonglong2 g = row2[cidx];
            const ulonglong2 t = tm2[cidx];
            const ulonglong2 m = m2[cidx];
            tp += __popcll(g.x & t.x) + __popcll(g.y & t.y);
            gp += __popcll(g.x & m.x) + __popcll(g.y & m.y);
            if (HASH) {
                const int w0 = 2 * cidx;
                if (w0 < A.Wn) {
                    const uint64_t x = g.x & m.x;
                    h0 += mix64(x + (uint64_t)(w0 + 1) * 0x9E3779B97F4A7C15ULL);
                    h1 += mix64((x ^ 0xD6E8FEB86659FD93ULL) + (uint64_t)(w0 + 1) * 0xC2B2AE3D27D4EB4FULL);
                }
                if (w0 + 1 < A.Wn) {
                    const uint64_t x = g.y & m.y;
                    h0 += mix64(x + (uint64_t)(w0 + 2) * 0x9E3779B97F4A7C15ULL);
                    h1 += mix64((x ^ 0xD6E8FEB86659FD93ULL) + (uint64_t)(w0 + 2) * 0xC2B2AE3D27D4EB4FULL);
                }
            }
        }
        __syncwarp();   // every lane has read the row: the buffer can be refilled
        if (lane == 0) {
            const int64_t nxt = g_idx + 2 * stride;
            if (nxt < A.G) {
                mbar_arrive_expect_tx(&my_bar[buf], row_bytes);
                tma_bulk_g2s(my_rows + (size_t)buf * A.W, A.genes + nxt * A.W, row_bytes, &my_bar[buf]);
            }
        }
        tp = __reduce_add_sync(0xffffffffu, tp);
        gp = __reduce_add_sync(0xffffffffu, gp);
        const int a = tp;              // tpgp
        const int c = gp - tp;         // tngp
        const int b = n_tp - tp;       // tpgn
        const int d = n_m - n_tp - c;  // tngn
        if (HASH) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                h0 += __shfl_xor_sync(0xffffffffu, h0, o);
                h1 += __shfl_xor_sync(0xffffffffu, h1, o);
            }
            if (lane == 0 && A.hash) {
                A.hash[g_idx * 2 + 0] = h0;
                A.hash[g_idx * 2 + 1] = h1;
            }
        }
        if (lane == 0 && A.counts)
            reinterpret_cast<int4 *>(A.counts)[g_idx] = make_int4(a, c, b, d);   // tpgp,tngp,tpgn,tngn
        if (A.p) {
            const double pv = fisher_two_sided_warp(lut, a, b, c, d, lane);
            if (lane == 0) A.p[g_idx] = pv;
        }
    }
}

#endif  // !SB_HOST_EMUL

}  // namespace sb
