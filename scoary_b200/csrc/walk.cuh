// walk.cuh -- K1 (pack into walk order), label shuffles, K4/K5 (PhyloTree walks).
//
// The pairwise-comparisons DP of class PhyloTree (scoary/classes.py:199-572)
// in max-plus form.  Per node and per state c in {AB, Ab, aB, ab, 0 = no free
// path} the reference keeps (max pairs, max pro-pairs | max pairs, max
// anti-pairs | max pairs) with -1 = unreachable.  Lexicographic maxima are
// plain integer maxima of the keys
//     kp = (pairs << SH) + pro,      ka = (pairs << SH) + anti      (SH: 2^SH > n_leaves/2)
// with unreachable = a large negative number, so a node update is
//     out[c] = max(L[c] + max_x R[x], max_x L[x] + R[c])                       c in AB,Ab,aB,ab
//     out[0] = max(L0 + R0, L_AB + R_ab + b+, L_ab + R_AB + b+, L_Ab + R_aB + b-, L_aB + R_Ab + b-)
// (b+ = 2^SH+1, b- = 2^SH in the pro pass; swapped in the anti pass), which is
// classes.py:268-457 and :459-572 with the "given max pairs" tie rule folded
// into the key.  The root takes three independent maxima (classes.py:246-249).
//
// Because the tree is the same for every (gene, labelling), it is compiled
// once on the host into a small stack program:
//   CHERRY        acc <- node of two leaves           (push the old acc first if PUSH)
//   LEAF  xN      acc <- combine(acc, next leaf)       N times
//   MERGE xN      acc <- combine(pop(), acc)           N times
// Children are ordered so the deeper side is evaluated first (Strahler
// order): the stack never exceeds log2(#cherries) entries, and leaves are
// consumed strictly left to right, so gene and label bits are read as a
// stream.  One thread = one (gene, labelling) walk; all threads of a block run
// the same program on the same labelling, so every branch is uniform.
#pragma once
#include "common.cuh"

namespace sb {

constexpr int WALK_THREADS = 128;
constexpr int WALK_NEG = -(1 << 30);   // unreachable; keys stay < 2^28 (n_leaves <= 32766), so inv+inv >= INT_MIN
constexpr int OP_CHERRY = 0, OP_LEAF = 1, OP_MERGE = 2, OP_CHERRY_PUSH = 3;
constexpr int PERMS_PER_BLOCK_MAX = 32;
constexpr uint32_t SHUFFLE_DOMAIN = 0x5C0A27u;

// ---------------------------------------------------------------- K1: gather + transpose
// genes [G][W] uint64 (isolate columns)  ->  genesT [W32p][Gs] uint32 where bit b
// of word w of gene g = gene g at the leaf consumed at walk position 32 w + b.
// Block = 32 genes x 8 word lanes; rows staged in shared memory with an odd
// pitch so the gather is bank-conflict free.
__global__ void __launch_bounds__(256) pack_walk_order_kernel(const uint64_t *__restrict__ genes, int64_t G, int W,
                                                              const int32_t *__restrict__ walk_col, int n_leaves,
                                                              int W32p, int64_t Gs, uint32_t *__restrict__ genesT)
{
    extern __shared__ uint32_t s_rows[];   // [32][pitch]
    const int pitch = 2 * W + 1;
    const int64_t g0 = (int64_t)blockIdx.x * 32;
    const uint32_t *g32 = reinterpret_cast<const uint32_t *>(genes);
    for (int i = threadIdx.x; i < 32 * 2 * W; i += 256) {
        int r = i / (2 * W), c = i - r * (2 * W);
        int64_t g = g0 + r;
        s_rows[r * pitch + c] = (g < G) ? g32[g * (int64_t)(2 * W) + c] : 0u;
    }
    __syncthreads();
    const int gl = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int64_t g = g0 + gl;
    for (int w = wl; w < W32p; w += 8) {
        uint32_t word = 0;
        const int base = w * 32;
#pragma unroll 4
        for (int b = 0; b < 32; ++b) {
            const int pos = base + b;
            if (pos < n_leaves) {
                const int col = __ldg(&walk_col[pos]);
                word |= ((s_rows[gl * pitch + (col >> 5)] >> (col & 31)) & 1u) << b;
            }
        }
        if (g < G) genesT[(int64_t)w * Gs + g] = word;
    }
}

// ---------------------------------------------------------------- label shuffles
// PermuteGTC (scoary/methods.py:1371-1384): random.shuffle of the trait labels
// = Fisher-Yates.  One thread = one permutation; its label bit-vector lives in
// shared memory ([word][thread], conflict free).  Same stream as the oracle's
// so_shuffle_labels: step s = n-1-i draws u64 from Philox block s/2, j = mulhi(u, i+1).
// Output: labelsW [P][W32p] in walk order (bit b of word w = label of the leaf
// consumed at position 32 w + b); optionally the labels by leaf id (test hook).
__global__ void __launch_bounds__(64) shuffle_labels_kernel(const uint32_t *__restrict__ labels_leaf /*[W32]*/,
                                                            int n_leaves, int W32, int W32p,
                                                            const int32_t *__restrict__ leaf_of_pos, uint64_t seed,
                                                            int trait, int P, uint32_t *__restrict__ labelsW,
                                                            uint8_t *__restrict__ dbg_leaf /* [P][n_leaves] or null */)
{
    extern __shared__ uint32_t s_lab[];   // [W32][64]
    const int T = 64, tid = threadIdx.x;
    const int perm = blockIdx.x * T + tid;
    for (int w = 0; w < W32; ++w) s_lab[w * T + tid] = labels_leaf[w];
    if (perm >= P) return;   // no block-wide sync below
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t rnd[4] = {0, 0, 0, 0};
    for (int i = n_leaves - 1; i >= 1; --i) {
        const int s = n_leaves - 1 - i;
        if ((s & 1) == 0) philox4x32_10((uint32_t)(s >> 1), (uint32_t)perm, (uint32_t)trait, SHUFFLE_DOMAIN, k0, k1, rnd);
        const uint64_t u = (s & 1) ? ((uint64_t)rnd[2] | ((uint64_t)rnd[3] << 32))
                                   : ((uint64_t)rnd[0] | ((uint64_t)rnd[1] << 32));
        const int j = (int)__umul64hi(u, (uint64_t)(i + 1));
        const uint32_t wi = s_lab[(i >> 5) * T + tid], wj = s_lab[(j >> 5) * T + tid];
        const uint32_t bi = (wi >> (i & 31)) & 1u, bj = (wj >> (j & 31)) & 1u;
        if (bi != bj) {
            s_lab[(i >> 5) * T + tid] ^= (1u << (i & 31));
            s_lab[(j >> 5) * T + tid] ^= (1u << (j & 31));
        }
    }
    for (int w = 0; w < W32p; ++w) {
        uint32_t word = 0;
        for (int b = 0; b < 32; ++b) {
            const int pos = w * 32 + b;
            if (pos < n_leaves) {
                const int leaf = __ldg(&leaf_of_pos[pos]);
                word |= ((s_lab[(leaf >> 5) * T + tid] >> (leaf & 31)) & 1u) << b;
            }
        }
        labelsW[(int64_t)perm * W32p + w] = word;
    }
    if (dbg_leaf) {
        for (int k = 0; k < n_leaves; ++k)
            dbg_leaf[(int64_t)perm * n_leaves + k] = (uint8_t)((s_lab[(k >> 5) * T + tid] >> (k & 31)) & 1u);
    }
}

// ---------------------------------------------------------------- the DP
// state index: 0 AB (g=1,t=1), 1 Ab (g=1,t=0), 2 aB (g=0,t=1), 3 ab (g=0,t=0), 4 no free path
struct WalkState {
    int p[5];   // pro keys
    int a[5];   // anti keys
};

__device__ __forceinline__ int max5(const int v[5])
{
    return __vimax3_s32(__vimax3_s32(v[0], v[1], v[2]), v[3], v[4]);
}

// node with two leaf children (classes.py:580-592 tips combined by :268-572)
__device__ __forceinline__ void walk_cherry(WalkState &o, int g1, int t1, int g2, int t2, int K)
{
    const int s1 = (1 - g1) * 2 + (1 - t1), s2 = (1 - g2) * 2 + (1 - t2);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int v = (s1 == c || s2 == c) ? 0 : WALK_NEG;
        o.p[c] = v;
        o.a[c] = v;
    }
    const bool comp = (s1 + s2) == 3;
    const bool propair = (s1 == 0) || (s1 == 3);       // AB/ab pair supports, Ab/aB pair opposes
    o.p[4] = comp ? (propair ? K + 1 : K) : WALK_NEG;
    o.a[4] = comp ? (propair ? K : K + 1) : WALK_NEG;
}

// acc <- combine(acc, leaf (g, t)); t is block-uniform, g is per thread
__device__ __forceinline__ void walk_leaf(WalkState &s, int g, int t, int K)
{
    const int Mp = max5(s.p), Ma = max5(s.a);
    const bool G1 = g != 0;
    if (t) {   // leaf is AB (g) or aB (!g); its complement is ab (pro pair) or Ab (anti pair)
        const int np4 = G1 ? s.p[3] + (K + 1) : s.p[1] + K;
        const int na4 = G1 ? s.a[3] + K : s.a[1] + (K + 1);
        s.p[0] = G1 ? Mp : s.p[0];
        s.p[2] = G1 ? s.p[2] : Mp;
        s.a[0] = G1 ? Ma : s.a[0];
        s.a[2] = G1 ? s.a[2] : Ma;
        s.p[4] = np4;
        s.a[4] = na4;
    } else {   // leaf is Ab (g) or ab (!g); complement aB (anti pair) or AB (pro pair)
        const int np4 = G1 ? s.p[2] + K : s.p[0] + (K + 1);
        const int na4 = G1 ? s.a[2] + (K + 1) : s.a[0] + K;
        s.p[1] = G1 ? Mp : s.p[1];
        s.p[3] = G1 ? s.p[3] : Mp;
        s.a[1] = G1 ? Ma : s.a[1];
        s.a[3] = G1 ? s.a[3] : Ma;
        s.p[4] = np4;
        s.a[4] = na4;
    }
}

__device__ __forceinline__ void merge_pass(const int L[5], const int R[5], int out[5], int bpro, int banti)
{
    const int ML = max5(L), MR = max5(R);
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = __viaddmax_s32(L[c], MR, ML + R[c]);
    const int nf = L[4] + R[4];
    const int pp = __viaddmax_s32(L[0], R[3], L[3] + R[0]) + bpro;
    const int ap = __viaddmax_s32(L[1], R[2], L[2] + R[1]) + banti;
    out[4] = max(__vimax3_s32(nf, pp, ap), WALK_NEG);
}

__device__ __forceinline__ void walk_merge(const WalkState &L, WalkState &acc, int K)
{
    WalkState o;
    merge_pass(L.p, acc.p, o.p, K + 1, K);
    merge_pass(L.a, acc.a, o.a, K, K + 1);
    acc = o;
}

struct WalkArgs {
    const uint32_t *genesT;    // [W32p][Gs]
    int64_t Gs;
    const int64_t *gene_idx;   // [S] or null (identity)
    int64_t S;
    const uint32_t *labelsW;   // [n_label_rows][W32p], walk order
    int32_t W32p;
    const uint16_t *ops;       // [n_ops]: (count << 2) | type
    int32_t n_ops;
    int32_t n_leaves;
    int32_t shift;             // SH
    int32_t stack_depth;       // max pushes
    int32_t P;                 // labellings (permute mode)
    int32_t perms_per_block;
    int32_t n_chunks;          // ceil(P / perms_per_block)
    const int32_t *unperm;     // [S][3] (permute mode): unpermuted Total, Pro, Anti
    int32_t *pairs;            // [S][3] (pairs mode output)
    uint32_t *hitbits;         // [S][n_chunks] (permute mode output)
};

// One full tree walk for this thread's gene under the labelling `lab` (shared
// memory, walk order).  Returns the root state in acc.
__device__ __forceinline__ void walk_tree(const WalkArgs &A, const uint32_t *__restrict__ gcol /* genesT + gene */,
                                          const uint32_t *lab, int *stk, WalkState &acc)
{
    const int K = 1 << A.shift;
    constexpr int T = WALK_THREADS;
    int pos = 0, sp = 0;
    uint32_t gw = 0, lw = 0;
    uint32_t gnext = __ldg(gcol);
    auto next_bits = [&](int &g, int &t) {
        if ((pos & 31) == 0) {
            const int w = pos >> 5;
            gw = gnext;
            lw = lab[w];
            if (w + 1 < A.W32p) gnext = __ldg(gcol + (int64_t)(w + 1) * A.Gs);
        }
        g = (int)(gw & 1u);
        t = (int)(lw & 1u);
        gw >>= 1;
        lw >>= 1;
        ++pos;
    };
    uint32_t op_next = __ldg(&A.ops[0]);
    for (int i = 0; i < A.n_ops; ++i) {
        const uint32_t op = op_next;
        if (i + 1 < A.n_ops) op_next = __ldg(&A.ops[i + 1]);
        const int type = op & 3, cnt = op >> 2;
        if (type == OP_LEAF) {
            for (int k = 0; k < cnt; ++k) {
                int g, t;
                next_bits(g, t);
                walk_leaf(acc, g, t, K);
            }
        } else if (type == OP_MERGE) {
            for (int k = 0; k < cnt; ++k) {
                --sp;
                WalkState L;
                const int *s = stk + sp * 10 * T;
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    L.p[c] = s[c * T];
                    L.a[c] = s[(5 + c) * T];
                }
                walk_merge(L, acc, K);
            }
        } else {
            if (type == OP_CHERRY_PUSH) {
                int *s = stk + sp * 10 * T;
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    s[c * T] = acc.p[c];
                    s[(5 + c) * T] = acc.a[c];
                }
                ++sp;
            }
            int g1, t1, g2, t2;
            next_bits(g1, t1);
            next_bits(g2, t2);
            walk_cherry(acc, g1, t1, g2, t2, K);
        }
    }
}

// root: three independent maxima (classes.py:246-249)
__device__ __forceinline__ void walk_root(const WalkState &s, int shift, int &total, int &pro, int &anti)
{
    const int mask = (1 << shift) - 1;
    total = max5(s.p) >> shift;
    pro = -1;
    anti = -1;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        if (s.p[c] >= 0) pro = max(pro, s.p[c] & mask);
        if (s.a[c] >= 0) anti = max(anti, s.a[c] & mask);
    }
}

// PERMUTE = false: K4, one labelling (labelsW row 0), writes pairs[S][3].
// PERMUTE = true : K5, grid.y = chunks of perms_per_block labellings; each block stages
//                  its chunk of label vectors with one TMA bulk copy and writes one
//                  32-bit word of hit flags per gene.
template <bool PERMUTE>
__global__ void __launch_bounds__(WALK_THREADS) walk_kernel(const WalkArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *s_lab = reinterpret_cast<uint32_t *>(smem_raw + 16);
    const int n_rows = PERMUTE ? A.perms_per_block : 1;
    int *stk = reinterpret_cast<int *>(s_lab + (size_t)n_rows * A.W32p) + threadIdx.x;

    const int chunk = PERMUTE ? blockIdx.y : 0;
    const int perm0 = chunk * A.perms_per_block;
    const int rows = PERMUTE ? min(A.perms_per_block, A.P - perm0) : 1;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        const uint32_t bytes = (uint32_t)rows * (uint32_t)A.W32p * 4u;
        mbar_arrive_expect_tx(bar, bytes);
        tma_bulk_g2s(s_lab, A.labelsW + (int64_t)perm0 * A.W32p, bytes, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);

    const int64_t s_idx = (int64_t)blockIdx.x * WALK_THREADS + threadIdx.x;
    const bool active = s_idx < A.S;
    const int64_t sc = active ? s_idx : (A.S - 1);   // inactive lanes redo the last gene (no divergence)
    const int64_t gene = A.gene_idx ? A.gene_idx[sc] : sc;
    const uint32_t *gcol = A.genesT + gene;

    if (!PERMUTE) {
        WalkState acc;
        walk_tree(A, gcol, s_lab, stk, acc);
        int total, pro, anti;
        walk_root(acc, A.shift, total, pro, anti);
        if (active) {
            A.pairs[s_idx * 3 + 0] = total;
            A.pairs[s_idx * 3 + 1] = pro;
            A.pairs[s_idx * 3 + 2] = anti;
        }
    } else {
        const long long u_total = A.unperm[sc * 3 + 0];
        const int u_pro = A.unperm[sc * 3 + 1], u_anti = A.unperm[sc * 3 + 2];
        const bool use_pro = u_pro >= u_anti;                 // methods.py:1333-1336
        const long long u_stat = use_pro ? u_pro : u_anti;
        uint32_t hits = 0;
        for (int r = 0; r < rows; ++r) {
            WalkState acc;
            walk_tree(A, gcol, s_lab + (size_t)r * A.W32p, stk, acc);
            int total, pro, anti;
            walk_root(acc, A.shift, total, pro, anti);
            const long long si = use_pro ? pro : anti;
            if (si * u_total >= u_stat * (long long)total) hits |= (1u << r);   // methods.py:1353-1355
        }
        if (active) A.hitbits[s_idx * A.n_chunks + chunk] = hits;
    }
}

// ---------------------------------------------------------------- hit-sequence reduction
// Permute's bookkeeping (methods.py:1348-1365) on the ordered hit flags.
__global__ void __launch_bounds__(256) reduce_hits_kernel(const uint32_t *__restrict__ hitbits, int64_t S, int n_chunks,
                                                          int perms_per_block, int P, int early_stop,
                                                          const int32_t *__restrict__ rmin, int32_t *__restrict__ r_out,
                                                          int32_t *__restrict__ n_done)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const uint32_t *h = hitbits + s * n_chunks;
    int r = 0, done = P;
    if (!early_stop) {
        for (int c = 0; c < n_chunks; ++c) {
            const int rows = min(perms_per_block, P - c * perms_per_block);
            const uint32_t m = rows >= 32 ? 0xffffffffu : ((1u << rows) - 1u);
            r += __popc(h[c] & m);
        }
    } else {
        bool stop = false;
        for (int c = 0; c < n_chunks && !stop; ++c) {
            const int rows = min(perms_per_block, P - c * perms_per_block);
            const uint32_t w = h[c];
            for (int b = 0; b < rows; ++b) {
                const int i = c * perms_per_block + b;
                r += (w >> b) & 1u;
                if (i >= 30 && r >= rmin[i]) {   // methods.py:1360-1363
                    done = i + 1;
                    stop = true;
                    break;
                }
            }
        }
    }
    r_out[s] = r;
    n_done[s] = done;
}

// ---------------------------------------------------------------- int32 pipe microbenchmark
// Dependent add/max chains, 8 independent chains per thread, as the walk DP issues them.
__global__ void __launch_bounds__(256) int32_peak_kernel(int *out, int iters, int seed)
{
    int x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = seed + threadIdx.x * (k + 1);
    const int a = seed | 1, b = -seed;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = __viaddmax_s32(x[k], a, b + k);
        }
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= x[k];
    if (s == 0x7fffffff) out[0] = s;
}

}  // namespace sb
