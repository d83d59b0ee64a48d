// fisher2.cuh -- EXPERIMENTAL second version of K2+K3 (built only with -DSB_FISHER_V2=1; the product
// library uses fisher.cuh).  Written after the last GPU session of round 1: its arithmetic is verified on the
// CPU through the host emulation (tests/test_host_emul.py), its speed is not measured yet.
//
// Why: ncu shows fisher_kernel bound by FP64 issue and latency, not by HBM (3.5 % of the HBM roofline at
// C3).  In version 1 a whole warp serves one gene and every lane evaluates 8 terms of the hypergeometric
// tail per round -- one exp and seven dependent divisions per lane, and typically only 1-2 rounds are
// needed, so most FP64 warp instructions carry few useful lanes.  Here
//   * a warp serves FOUR genes (8 lanes each): every FP64 warp instruction does work for four tables;
//   * a lane evaluates a block of 32 consecutive terms: ONE exp for the block's first term, the other
//     31 through the ratio recurrence written as a backward Horner scheme with a common denominator,
//         sum_{j<m} prod_{i<=j} num_i/den_i = U/V,  U <- V*den_i + num_i*U,  V <- V*den_i  (i = m-1 .. 1),
//     i.e. three multiplications and one FMA per term and ONE division per block (products stay below
//     1e271 for N <= 32766);
//   * the search for the first included term on the far side of the mode is 8-ary per group.
// The p-value rule, the canonical table orientation (bit-identical p for symmetric variants), the tie
// tolerance and the 2^-80 cut are those of fisher.cuh.
#pragma once
#include "fisher.cuh"

#ifdef SB_HOST_EMUL
#define SB_ANY(pred) sb_emul_any(pred)
#define SB_FMA(a, b, c) fma(a, b, c)
#else
#define SB_ANY(pred) __any_sync(0xffffffffu, pred)
#define SB_FMA(a, b, c) __fma_rn(a, b, c)
#endif

namespace sb {

constexpr int F2_LANES = 8;     // lanes per gene
constexpr int F2_GENES = 4;     // genes per warp
constexpr int F2_BLOCK = 32;    // terms per lane per round

SB_DEV double group_sum(double v)
{
    v += SB_SHFL_XOR(v, 4);
    v += SB_SHFL_XOR(v, 2);
    v += SB_SHFL_XOR(v, 1);
    return v;
}

// Per group: sum pmf(x) for x = x0, x0+dir, ..., count terms (non-increasing); count = 0 for a group with
// nothing to add.  All 32 lanes call this together (the collectives are warp-wide).
SB_DEV double fisher2_tail(const double2 *lut, int x0, int dir, int count, int n1, int n2, int n, dd logp_a, dd S_a,
                           double pexact, int l8, int lane0)
{
    double acc = 0.0;
    const double cut = pexact * 8.271806125530277e-25;   // 2^-80
    bool live = count > 0;
    for (int base = 0; SB_ANY(live && base < count); base += F2_LANES * F2_BLOCK) {
        const int kb = base + l8 * F2_BLOCK;
        double t = 0.0;
        if (live && kb < count) {
            const int x = x0 + dir * kb;
            const int m = min(F2_BLOCK, count - kb);
            dd dl = dd_sub(S_a, fisher_S(lut, x, n1, n2, n));   // log pmf(x) - log pmf(a)
            dd L = dd_add(logp_a, dl);
            t = exp(L.hi) * (1.0 + L.lo);
            double A, B, C, D;      // pmf(next)/pmf(cur) = (A * B) / (C * D); A, B step down, C, D step up
            if (dir > 0) { A = (double)(n1 - x); B = (double)(n - x); C = (double)(x + 1); D = (double)(n2 - n + x + 1); }
            else { A = (double)x; B = (double)(n2 - n + x); C = (double)(n1 - x + 1); D = (double)(n - x + 1); }
            double U = 1.0, V = 1.0;
            for (int i = m - 1; i >= 1; --i) {
                const double k = (double)(i - 1);
                const double num = SB_DMUL(A - k, B - k), den = SB_DMUL(C + k, D + k);   // exact: integers < 2^53
                const double Vd = SB_DMUL(V, den);
                U = SB_FMA(num, U, Vd);
                V = Vd;
            }
            acc += SB_DMUL(t, SB_DDIV(U, V));
        }
        const double first = SB_SHFL(t, lane0);           // the group's lane 0: first term of the round
        if (first < cut) live = false;
    }
    return group_sum(acc);
}

// Two-sided Fisher exact p for four tables at once: lanes 8g..8g+7 hold table g ([[a, b], [c, d]]);
// valid = false for a group without a table (it still takes part in the collectives).
SB_DEV double fisher2_two_sided(const double2 *lut, int a, int b, int c, int d, bool valid, int lane)
{
    const int l8 = lane & (F2_LANES - 1), lane0 = lane & ~(F2_LANES - 1);
    if (!valid) { a = 0; b = 0; c = 0; d = 0; }           // an empty table only ever touches lut[0]
    bool done = a + b == 0 || c + d == 0 || a + c == 0 || b + d == 0;
    {   // canonical orientation (see fisher.cuh)
        int d0 = min(a, d), d1 = max(a, d), o0 = min(b, c), o1 = max(b, c);
        if (o0 < d0 || (o0 == d0 && o1 < d1)) {
            int t0 = d0, t1 = d1;
            d0 = o0; d1 = o1; o0 = t0; o1 = t1;
        }
        a = d0; d = d1; b = o0; c = o1;
    }
    const int n1 = a + b, n2 = c + d, n = a + c, M = n1 + n2;
    const int lo = max(0, n - n2), hi = min(n, n1);
    const int mode = (int)((double)((long long)(n + 1) * (long long)(n1 + 1)) / (double)(M + 2));
    if (a == mode) done = true;
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
        if (fabs(dm.hi + dm.lo) <= FISHER_TIE_TOL) done = true;
    }
    const int dir_obs = (a < mode) ? -1 : +1;
    const int cnt_obs = (a < mode) ? (a - lo + 1) : (hi - a + 1);
    double p = fisher2_tail(lut, a, dir_obs, done ? 0 : cnt_obs, n1, n2, n, logp_a, S_a, pexact, l8, lane0);

    const int dir2 = -dir_obs;
    const int K = (dir2 > 0) ? (hi - mode) : (mode - lo);
    int lo_k = 1, hi_k = done ? 1 : K + 1;
    while (SB_ANY(hi_k > lo_k)) {
        const bool act = hi_k > lo_k;
        const int len = act ? hi_k - lo_k : 1;
        const int stride = (len + F2_LANES - 1) >> 3;
        const int k = lo_k + l8 * stride;
        bool pred = false;
        if (act && k < hi_k) {
            dd dk = dd_sub(S_a, fisher_S(lut, mode + dir2 * k, n1, n2, n));
            pred = (dk.hi + dk.lo) <= FISHER_TIE_TOL;
        }
        const unsigned ball = (SB_BALLOT(pred) >> lane0) & 0xFFu;
        if (act) {
            if (ball == 0u) {
                lo_k = lo_k + ((len - 1) / stride) * stride + 1;
            } else {
                const int f = SB_FFS(ball) - 1;
                hi_k = lo_k + f * stride;
                lo_k = (f == 0) ? hi_k : (lo_k + (f - 1) * stride + 1);
            }
        }
    }
    const int kstar = lo_k;
    const int cnt2 = (done || kstar > K) ? 0 : K - kstar + 1;
    p += fisher2_tail(lut, mode + dir2 * kstar, dir2, cnt2, n1, n2, n, logp_a, S_a, pexact, l8, lane0);
    return done ? 1.0 : fmin(p, 1.0);
}

#ifndef SB_HOST_EMUL
#ifndef SB_FISHER2_THREADS
#define SB_FISHER2_THREADS 512
#endif
constexpr int FISHER2_THREADS = SB_FISHER2_THREADS;

// Same pipeline as fisher_kernel, four consecutive rows per warp and stage (one bulk copy).
template <bool LUT_SMEM, bool HASH>
__global__ void __launch_bounds__(FISHER2_THREADS) fisher2_kernel(const FisherArgs A)
{
    // Shared memory: [2 mbarriers per warp] [value & mask] [mask] [2 x 4 row buffers per warp] [LUT]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NW = FISHER2_THREADS / 32;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *s_tm = reinterpret_cast<uint64_t *>(smem_raw + 16 * NW);
    uint64_t *s_m = s_tm + A.W;
    uint64_t *s_rows = s_m + A.W;                                               // [NW][2][4][W]
    double2 *s_lut = reinterpret_cast<double2 *>(s_rows + (size_t)NW * 2 * F2_GENES * A.W);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 3, l8 = lane & 7;
    uint64_t *my_bar = bars + warp * 2;
    uint64_t *my_rows = s_rows + (size_t)warp * 2 * F2_GENES * A.W;
    const uint32_t row_bytes = (uint32_t)A.W * 8u;
    const int64_t n_batches = (A.G + F2_GENES - 1) / F2_GENES;
    const int64_t stride = (int64_t)gridDim.x * NW;
    const int64_t batch_first = (int64_t)blockIdx.x * NW + warp;

    auto issue = [&](int64_t batch, int buf) {
        const int64_t r0 = batch * F2_GENES;
        const uint32_t rows = (uint32_t)min((int64_t)F2_GENES, A.G - r0);
        mbar_arrive_expect_tx(&my_bar[buf], rows * row_bytes);
        tma_bulk_g2s(my_rows + (size_t)buf * F2_GENES * A.W, A.genes + r0 * A.W, rows * row_bytes, &my_bar[buf]);
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
    for (int w = tid; w < A.W; w += FISHER2_THREADS) {
        uint64_t m = A.tmask[w];
        s_m[w] = m;
        s_tm[w] = A.tvalue[w] & m;
    }
    if (LUT_SMEM) {
        for (int k = tid; k <= A.lut_n; k += FISHER2_THREADS) s_lut[k] = A.lut[k];
    }
    __syncthreads();
    const double2 *lut = LUT_SMEM ? s_lut : A.lut;

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
    for (int64_t bt = batch_first; bt < n_batches; bt += stride, ++it) {
        const int buf = it & 1;
        mbar_wait(&my_bar[buf], (uint32_t)((it >> 1) & 1));
        const int64_t g_idx = bt * F2_GENES + grp;
        const bool valid = g_idx < A.G;
        const ulonglong2 *row2 =
            reinterpret_cast<const ulonglong2 *>(my_rows + ((size_t)buf * F2_GENES + grp) * A.W);
        int tp = 0, gp = 0;
        uint64_t h0 = 0, h1 = 0;
        if (valid) {
            for (int cidx = l8; cidx < W2; cidx += F2_LANES) {
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
        }
        __syncwarp();   // every lane has read its row: the buffer can be refilled
        if (lane == 0) {
            const int64_t nxt = bt + 2 * stride;
            if (nxt < n_batches) issue(nxt, buf);
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
        const int a = tp;              // tpgp
        const int c = gp - tp;         // tngp
        const int b = n_tp - tp;       // tpgn
        const int d = n_m - n_tp - c;  // tngn
        if (valid && l8 == 0) {
            if (HASH && A.hash) {
                A.hash[g_idx * 2 + 0] = h0;
                A.hash[g_idx * 2 + 1] = h1;
            }
            if (A.counts) reinterpret_cast<int4 *>(A.counts)[g_idx] = make_int4(a, c, b, d);   // tpgp,tngp,tpgn,tngn
        }
        if (A.p) {
            const double pv = fisher2_two_sided(lut, a, b, c, d, valid, lane);
            if (valid && l8 == 0) A.p[g_idx] = pv;
        }
    }
}
#endif  // !SB_HOST_EMUL

}  // namespace sb
