// fisher.cuh -- K2+K3: per-gene 2x2 contingency tables by popcount over the packed bitset rows, then the
// two-sided Fisher exact p from a double-double log-factorial LUT -- for up to FISHER_MAX_TRAITS traits per row read.
//
// Replaces, per (gene, trait): Perform_statistics (scoary/methods.py:930-982) and the ss.fisher_exact call
// (methods.py:842-857); the reference loops traits outside genes (methods.py:771, :791), here every gene row is read
// once for all traits of the launch.  The p-value rule is SciPy 1.18.1's (scipy/stats/_stats_py.py, fisher_exact,
// two-sided):
//   p = sum_{x on the observed side, from the extreme up to a} pmf(x)
//     + sum_{x on the other side of the mode with pmf(x) <= pmf(a)(1+1e-14)} pmf(x)
//   p = 1 when pmf(a) ~= pmf(mode);  p = min(p, 1).
//
// Data movement: a persistent grid; every warp streams its own batches of FOUR consecutive gene rows into shared
// memory with 1-D TMA bulk copies (cp.async.bulk + its own pair of mbarriers, double buffered; no block-wide barrier
// in the loop).  Eight lanes own one row: 128-bit shared loads, __popcll, three shuffles.
//
// Arithmetic (FP64-issue bound, so the design minimises FP64 instructions per table):
//   * a warp serves four tables (8 lanes each): every FP64 warp instruction does work for four tables;
//   * a lane evaluates a block of consecutive hypergeometric terms from ONE exp: the first term from the
//     double-double LUT, the rest through the ratio recurrence
//         pmf(x+1)/pmf(x) = (n1-x)(n-x) / ((x+1)(n2-n+x+1))
//     written as a backward Horner scheme with a common denominator,
//         sum_{j<m} prod_{i<=j} num_i/den_i = U/V,  U <- V*den_i + num_i*U,  V <- V*den_i  (i = m-1 .. 1),
//     three multiplications and one FMA per term, ONE division per block (products of <= 31 factors stay
//     below 1e280 for N <= 32766);
//   * the first included term on the far side of the mode is looked for in a window around the reflection of
//     `a` about the mode (one round of 8 candidates per table), with the 8-ary search as the fallback;
//   * tables whose `a` lies within 2.5 standard deviations of the mode -- almost every null gene -- are summed
//     the short way round: p = 1 - sum of the terms strictly between `a` and the far-side boundary (a few dozen
//     terms instead of the several hundred the two tails hold above 2^-64 pmf(a)).  The absolute error of that
//     sum is ~1e-15 and the path is only taken for p >= 0.01 (checked on the result; otherwise the tails are
//     summed), so the relative error stays below 1e-12.
// Symmetric variants of a table (row swap, column swap, transpose) are put in one canonical orientation first and
// therefore give bit-identical p; a table's p does not depend on the tables that share its warp.
#pragma once
#include "common.cuh"

// Warp collectives and rounding intrinsics of the Fisher code: CUDA spellings in a device build, the emulation
// driver's versions (32 host threads in lockstep) under SB_HOST_EMUL.
#ifdef SB_HOST_EMUL
#define SB_SHFL_XOR(v, o) sb_emul_shfl_xor(v, o)
#define SB_SHFL(v, src) sb_emul_shfl(v, src)
#define SB_BALLOT(pred) sb_emul_ballot(pred)
#define SB_ANY(pred) sb_emul_any(pred)
#define SB_FFS(x) __builtin_ffs((int)(x))
#define SB_DMUL(a, b) sb_emul_mul(a, b)
#define SB_DDIV(a, b) sb_emul_div(a, b)
#define SB_FMA(a, b, c) fma(a, b, c)
#else
#define SB_SHFL_XOR(v, o) __shfl_xor_sync(0xffffffffu, v, o)
#define SB_SHFL(v, src) __shfl_sync(0xffffffffu, v, src)
#define SB_BALLOT(pred) __ballot_sync(0xffffffffu, pred)
#define SB_ANY(pred) __any_sync(0xffffffffu, pred)
#define SB_FFS(x) __ffs(x)
#define SB_DMUL(a, b) __dmul_rn(a, b)
#define SB_DDIV(a, b) __ddiv_rn(a, b)
#define SB_FMA(a, b, c) __fma_rn(a, b, c)
#endif

namespace sb {

#ifndef SB_FISHER_THREADS
#define SB_FISHER_THREADS 768
#endif
constexpr int FISHER_THREADS = SB_FISHER_THREADS;
constexpr int FISHER_MAX_TRAITS = 8;       // traits per launch (their vectors are staged in shared memory)
constexpr double FISHER_TIE_TOL = 1e-12;   // |log pmf ratio| below this is a tie (exact ties give 0)
constexpr int F_LANES = 8;                 // lanes per table
constexpr int F_GENES = 4;                 // tables per warp
constexpr int F_BLOCK = 32;                // most terms a lane evaluates from one exp
constexpr double F_TAIL_CUT = 5.421010862427522e-20;   // 2^-64: a tail ends where its terms fall below this x pmf(a)
constexpr double F_NEAR_Z2 = 6.25;         // (a - mode)^2 <= this x variance: sum the complement instead of the tails
constexpr double F_NEAR_PMIN = 0.01;       // ... and keep that result only if p >= this

struct FisherArgs {
    const uint64_t *genes;   // [G][W]
    int64_t G;
    int32_t W;               // words per row (even)
    int32_t Wn;              // ceil(N/64): words that carry isolates (hash domain)
    const uint64_t *traits;  // [n_traits][2][W]: value, mask
    int32_t n_traits;
    const double2 *lut;      // [lut_n + 1] log k! as (hi, lo)
    int32_t lut_n;
    int32_t *counts;         // [n_traits][G][4] or null
    double *p;               // [n_traits][G] or null
    uint64_t *hash;          // [n_traits][G][2] or null
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

SB_DEV double group_sum(double v)
{
    v += SB_SHFL_XOR(v, 4);
    v += SB_SHFL_XOR(v, 2);
    v += SB_SHFL_XOR(v, 1);
    return v;
}

// Per table (group of 8 lanes): sum pmf(x) for x = x0, x0+dir, ..., `count` terms; count = 0 for a group with nothing
// to add.  The terms are dealt to the lanes in blocks of `blk` (<= F_BLOCK) consecutive terms.  TAIL: the terms are
// non-increasing and the sum stops once a round starts below cut.  All 32 lanes call this together (the collectives
// are warp-wide; TAIL is the same for the whole warp).
template <bool TAIL>
SB_DEV double fisher_sum(const double2 *lut, int x0, int dir, int count, int blk, int n1, int n2, int n, dd logp_a,
                         dd S_a, double cut, int l8, int lane0)
{
    double acc = 0.0;
    bool live = count > 0;
    for (int base = 0; SB_ANY(live && base < count); base += F_LANES * blk) {
        const int kb = base + l8 * blk;
        double t = 0.0;
        if (live && kb < count) {
            const int x = x0 + dir * kb;
            const int m = min(blk, count - kb);
            dd dl = dd_sub(S_a, fisher_S(lut, x, n1, n2, n));   // log pmf(x) - log pmf(a)
            dd L = dd_add(logp_a, dl);
            t = exp(L.hi) * (1.0 + L.lo);
            // pmf(next)/pmf(cur) = (A * B) / (C * D); A, B step down, C, D step up.  Factors of the LAST ratio used:
            double A, B, C, D;
            const double k = (double)(m - 2);
            if (dir > 0) { A = (double)(n1 - x) - k; B = (double)(n - x) - k; C = (double)(x + 1) + k; D = (double)(n2 - n + x + 1) + k; }
            else { A = (double)x - k; B = (double)(n2 - n + x) - k; C = (double)(n1 - x + 1) + k; D = (double)(n - x + 1) + k; }
            double U = 1.0, V = 1.0;
            for (int i = m - 1; i >= 1; --i) {
                const double num = SB_DMUL(A, B), den = SB_DMUL(C, D);   // exact: integers < 2^53
                const double Vd = SB_DMUL(V, den);
                U = SB_FMA(num, U, Vd);
                V = Vd;
                A += 1.0; B += 1.0; C -= 1.0; D -= 1.0;
            }
            acc += SB_DMUL(t, SB_DDIV(U, V));
        }
        if (TAIL) {
            const double first = SB_SHFL(t, lane0);           // the group's lane 0: first term of the round
            if (first < cut) live = false;
        }
    }
    return group_sum(acc);
}

// Two-sided Fisher exact p for four tables at once: lanes 8g..8g+7 hold table g ([[a, b], [c, d]]);
// valid = false for a group without a table (it still takes part in the collectives).
SB_DEV double fisher_two_sided(const double2 *lut, int a, int b, int c, int d, bool valid, int lane)
{
    const int l8 = lane & (F_LANES - 1), lane0 = lane & ~(F_LANES - 1);
    if (!valid) { a = 0; b = 0; c = 0; d = 0; }           // an empty table only ever touches lut[0]
    bool done = a + b == 0 || c + d == 0 || a + c == 0 || b + d == 0;
    // The p-value is invariant under swapping rows, swapping columns and transposing.  Put the table in a canonical
    // orientation first so that all eight variants run the very same arithmetic and return bit-identical p (ties
    // stay ties for the BH tie rule and the p-sort).  Invariant: the unordered pair of unordered pairs {{a,d},{b,c}}.
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
    // SciPy: mode = int((n + 1)(n1 + 1) / (M + 2)) in double; the operands are integers < 2^31 and a quotient that is
    // not an integer is at least 1 / (M + 2) away from one, so the unsigned division gives the same floor
    const int mode = (int)(((unsigned)(n + 1) * (unsigned)(n1 + 1)) / (unsigned)(M + 2));
    if (a == mode) done = true;
    // log pmf(a)
    dd base = dd_make(lut[n1]);
    base = dd_add(base, dd_make(lut[n2]));
    base = dd_add(base, dd_make(lut[n]));
    base = dd_add(base, dd_make(lut[M - n]));
    base = dd_sub(base, dd_make(lut[M]));
    const dd S_a = fisher_S(lut, a, n1, n2, n);
    const dd logp_a = dd_sub(base, S_a);
    const double pexact = exp(logp_a.hi) * (1.0 + logp_a.lo);
    const int dist = (a < mode) ? (mode - a) : (a - mode);
    if (dist == 1) {   // pexact ~= pmode -> 1: the pmf is strictly unimodal, so only a neighbour of the mode can tie with it
        dd dm = dd_sub(S_a, fisher_S(lut, mode, n1, n2, n));
        if (fabs(dm.hi + dm.lo) <= FISHER_TIE_TOL) done = true;
    }
    const int dir_obs = (a < mode) ? -1 : +1;          // away from the mode on the observed side
    const int cnt_obs = (a < mode) ? (a - lo + 1) : (hi - a + 1);

    // Other side: y_k = mode + dir2 * k, k = 1..K; included iff log pmf(y_k) - log pmf(a) <= tol.  pmf decreases
    // with k, so the included terms are k >= kstar.  Invariant of the search: kstar in [lo_k, hi_k], and hi_k is
    // either K + 1 or a k known to be included.  First round: a window of 8 around the reflection of a.
    const int dir2 = -dir_obs;
    const int K = (dir2 > 0) ? (hi - mode) : (mode - lo);
    int lo_k = 1, hi_k = done ? 1 : K + 1;
    bool first_round = true;
    while (SB_ANY(hi_k > lo_k)) {
        const bool act = hi_k > lo_k;
        const int len = act ? hi_k - lo_k : 1;
        int stride = (len + F_LANES - 1) >> 3, k0 = lo_k;
        if (first_round && len > F_LANES) {
            stride = 1;
            k0 = min(max(dist - 3, 1), hi_k - F_LANES);
        }
        const int k = k0 + l8 * stride;
        bool pred = false;
        if (act && k < hi_k) {
            dd dk = dd_sub(S_a, fisher_S(lut, mode + dir2 * k, n1, n2, n));
            pred = (dk.hi + dk.lo) <= FISHER_TIE_TOL;
        }
        const unsigned ball = (SB_BALLOT(pred) >> lane0) & 0xFFu;
        if (act) {
            if (ball == 0u) {                       // every candidate below hi_k is excluded
                const int last = k0 + ((hi_k - 1 - k0) / stride < F_LANES - 1 ? (hi_k - 1 - k0) / stride : F_LANES - 1) * stride;
                lo_k = last + 1;
            } else {
                const int f = SB_FFS(ball) - 1;
                hi_k = k0 + f * stride;
                if (f > 0) lo_k = k0 + (f - 1) * stride + 1;     // f == 0: everything from lo_k up to k0 is still open
            }
        }
        first_round = false;
    }
    const int kstar = lo_k;

    // Near the mode: 1 - (the terms strictly between a and y_kstar), x = a + dir2 * j, j = 1 .. cc
    // (a - mode)^2 <= F_NEAR_Z2 x variance, variance = n n1 n2 (M - n) / (M^2 (M - 1)), without the divisions
    const double dM = (double)M;
    const double lhs = ((double)dist * (double)dist) * (dM * dM) * (double)max(M - 1, 1);
    const double rhs = F_NEAR_Z2 * (((double)n * (double)n1) * ((double)n2 * (double)(M - n)));
    const int cc = dist + kstar - 1;
    const bool near_mode = !done && lhs <= rhs && cc <= F_LANES * F_BLOCK;
    const int blk_c = min(F_BLOCK, (cc + F_LANES - 1) >> 3);
    const double inner = fisher_sum<false>(lut, a + dir2, dir2, near_mode ? cc : 0, max(blk_c, 1), n1, n2, n, logp_a, S_a, 0.0,
                                    l8, lane0);
    const double p_near = 1.0 - inner;
    const bool tails = !done && !(near_mode && p_near >= F_NEAR_PMIN);

    double p = 0.0;
    if (SB_ANY(tails)) {       // most warps hold four null genes: no tail to sum
        const double cut = pexact * F_TAIL_CUT;
        p = fisher_sum<true>(lut, a, dir_obs, tails ? cnt_obs : 0, F_BLOCK, n1, n2, n, logp_a, S_a, cut, l8, lane0);
        const int cnt2 = (!tails || kstar > K) ? 0 : K - kstar + 1;
        p += fisher_sum<true>(lut, mode + dir2 * kstar, dir2, cnt2, F_BLOCK, n1, n2, n, logp_a, S_a, cut, l8, lane0);
    }
    return done ? 1.0 : (tails ? fmin(p, 1.0) : p_near);
}

#ifndef SB_HOST_EMUL   // the kernel itself (TMA pipeline, popcounts) is device-only
template <bool LUT_SMEM, bool HASH>
__global__ void __launch_bounds__(FISHER_THREADS) fisher_kernel(const FisherArgs A)
{
    // Shared memory: [2 mbarriers per warp] [per trait: value & mask, mask] [2 x 4 row buffers per warp] [LUT]
    // The block may be launched with fewer than FISHER_THREADS threads: long rows (N > ~7 000 isolates) leave
    // room for fewer warps' row buffers.
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NW = (int)blockDim.x >> 5, NT = (int)blockDim.x;
    __shared__ int s_tot[FISHER_MAX_TRAITS][2];      // per trait: positives, non-missing
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *s_tr = reinterpret_cast<uint64_t *>(smem_raw + 16 * NW);          // [n_traits][2][W]: value & mask, mask
    uint64_t *s_rows = s_tr + (size_t)A.n_traits * 2 * A.W;                     // [NW][2][4][W]
    double2 *s_lut = reinterpret_cast<double2 *>(s_rows + (size_t)NW * 2 * F_GENES * A.W);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 3, l8 = lane & 7;
    uint64_t *my_bar = bars + warp * 2;
    uint64_t *my_rows = s_rows + (size_t)warp * 2 * F_GENES * A.W;
    const uint32_t row_bytes = (uint32_t)A.W * 8u;
    const int64_t n_batches = (A.G + F_GENES - 1) / F_GENES;
    const int64_t stride = (int64_t)gridDim.x * NW;
    const int64_t batch_first = (int64_t)blockIdx.x * NW + warp;
    const int nT = A.n_traits;

    auto issue = [&](int64_t batch, int buf) {
        const int64_t r0 = batch * F_GENES;
        const uint32_t rows = (uint32_t)min((int64_t)F_GENES, A.G - r0);
        mbar_arrive_expect_tx(&my_bar[buf], rows * row_bytes);
        tma_bulk_g2s(my_rows + (size_t)buf * F_GENES * A.W, A.genes + r0 * A.W, rows * row_bytes, &my_bar[buf]);
    };
    if (lane == 0) {
        mbar_init(&my_bar[0], 1);
        mbar_init(&my_bar[1], 1);
        fence_barrier_init();
#pragma unroll
        for (int bf = 0; bf < 2; ++bf) {
            const int64_t bt = batch_first + bf * stride;
            if (bt < n_batches) issue(bt, bf);
        }
    }
    for (int i = tid; i < nT * A.W; i += NT) {
        const int t = i / A.W, w = i - t * A.W;
        const uint64_t m = A.traits[((size_t)t * 2 + 1) * A.W + w];
        s_tr[((size_t)t * 2 + 1) * A.W + w] = m;
        s_tr[((size_t)t * 2) * A.W + w] = A.traits[((size_t)t * 2) * A.W + w] & m;
    }
    if (LUT_SMEM) {
        for (int k = tid; k <= A.lut_n; k += NT) s_lut[k] = A.lut[k];
    }
    __syncthreads();
    for (int t = warp; t < nT; t += NW) {       // trait totals, one warp per trait
        int n_tp = 0, n_m = 0;
        for (int w = lane; w < A.W; w += 32) {
            n_tp += __popcll(s_tr[((size_t)t * 2) * A.W + w]);
            n_m += __popcll(s_tr[((size_t)t * 2 + 1) * A.W + w]);
        }
        n_tp = __reduce_add_sync(0xffffffffu, n_tp);
        n_m = __reduce_add_sync(0xffffffffu, n_m);
        if (lane == 0) { s_tot[t][0] = n_tp; s_tot[t][1] = n_m; }
    }
    __syncthreads();   // trait vectors, totals and LUT staged (the only block-wide barriers)
    const double2 *lut = LUT_SMEM ? s_lut : A.lut;
    const int W2 = A.W >> 1;

    int it = 0;
    for (int64_t bt = batch_first; bt < n_batches; bt += stride, ++it) {
        const int buf = it & 1;
        mbar_wait(&my_bar[buf], (uint32_t)((it >> 1) & 1));
        const int64_t g_idx = bt * F_GENES + grp;
        const bool valid = g_idx < A.G;
        const ulonglong2 *row2 =
            reinterpret_cast<const ulonglong2 *>(my_rows + ((size_t)buf * F_GENES + grp) * A.W);
        // One trait at a time (not unrolled: fisher_two_sided is large).  The row buffer is handed back to the TMA
        // pipeline after the LAST trait's popcount; the batch after next is what it is refilled with, so the copy
        // still has a whole batch of Fisher arithmetic to land in.
#pragma unroll 1
        for (int t = 0; t < nT; ++t) {
            const ulonglong2 *tm2 = reinterpret_cast<const ulonglong2 *>(s_tr + ((size_t)t * 2) * A.W);
            const ulonglong2 *m2 = reinterpret_cast<const ulonglong2 *>(s_tr + ((size_t)t * 2 + 1) * A.W);
            int tp = 0, gp = 0;
            uint64_t h0 = 0, h1 = 0;
            if (valid) {
                for (int cidx = l8; cidx < W2; cidx += F_LANES) {
                    const ulonglong2 g = row2[cidx];
                    const ulonglong2 tv = tm2[cidx];
                    const ulonglong2 m = m2[cidx];
                    tp += __popcll(g.x & tv.x) + __popcll(g.y & tv.y);
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
            }
            if (t == nT - 1) {
                __syncwarp();   // every lane has read its row for every trait: the buffer can be refilled
                if (lane == 0) {
                    const int64_t nxt = bt + 2 * stride;
                    if (nxt < n_batches) issue(nxt, buf);
                }
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                tp += __shfl_xor_sync(0xffffffffu, tp, o);
                gp += __shfl_xor_sync(0xffffffffu, gp, o);
                if (HASH) {
                    h0 += __shfl_xor_sync(0xffffffffu, h0, o);
                    h1 += __shfl_xor_sync(0xffffffffu, h1, o);
                }
            }
            const int n_tp = s_tot[t][0], n_m = s_tot[t][1];
            const int a = tp;              // tpgp
            const int c = gp - tp;         // tngp
            const int b = n_tp - tp;       // tpgn
            const int d = n_m - n_tp - c;  // tngn
            if (valid && l8 == 0) {
                if (HASH && A.hash) {
                    A.hash[((int64_t)t * A.G + g_idx) * 2 + 0] = h0;
                    A.hash[((int64_t)t * A.G + g_idx) * 2 + 1] = h1;
                }
                if (A.counts) reinterpret_cast<int4 *>(A.counts)[(int64_t)t * A.G + g_idx] = make_int4(a, c, b, d);   // tpgp,tngp,tpgn,tngn
            }
            if (A.p) {
                const double pv = fisher_two_sided(lut, a, b, c, d, valid, lane);
                if (valid && l8 == 0) A.p[(int64_t)t * A.G + g_idx] = pv;
            }
        }
    }
}

#endif  // !SB_HOST_EMUL

}  // namespace sb
