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
//   CHERRY_A / CHERRY_B   A (or B) <- node of two leaves
//   LEAF_A xN / LEAF_B xN A (or B) <- combine(A (or B), next leaf)      N times
//   MERGE_AB              A <- combine(A, B)
//   PUSH                  spill A to the DP stack (packed entries: shared memory; 32-bit entries: local memory)
//   MERGE_POP             A <- combine(pop(), A)          (xN for the 32-bit forms)
// (each in a packed 16-bit form for subtrees of <= 127 leaves -- two genes per register,
// .S16x2 DPX instructions -- and a 32-bit form above that, with WIDEN steps in between)
// A and B are two register-resident accumulators: a second child that is a
// "caterpillar" (a cherry plus single leaves) is evaluated in B while A stays
// live, so only second children with real branching touch the shared-memory
// stack.  Children are ordered to minimise the stack (<= log2(#cherries)
// entries), and leaves are consumed strictly left to right, so gene and label
// bits are read as a stream.  One thread = one (gene, labelling) walk; all
// threads of a block run the same program on the same labelling, so every
// branch is uniform.
#pragma once
#include "common.cuh"

namespace sb {

// The per-thread walk code is also compiled for the HOST by tests/host_emul (nvcc -DSB_HOST_EMUL, one
// simulated thread at a time), so the CPU test tier runs the very same DP arithmetic against the oracle.
// In a device build the macros below expand to exactly the CUDA spellings they stand for.
#ifdef SB_HOST_EMUL
struct EmulIdx { int x, y, z; };
static thread_local EmulIdx threadIdx, blockIdx;        // set by the emulation driver before each call
static thread_local int *sb_emul_shared = nullptr;      // the block's dynamic shared memory
#define SB_CONST static
#define SB_LDG(p) (*(p))
#define SB_KERNEL(bounds) inline void
#define SB_SHARED_STACK(name) int *name = sb_emul_shared
static inline unsigned sb_emul_vadd2(unsigned a, unsigned b)
{
    return (((a & 0xFFFFu) + (b & 0xFFFFu)) & 0xFFFFu) | (((a >> 16) + (b >> 16)) << 16);
}
static inline unsigned sb_emul_vmaxs2(unsigned a, unsigned b)
{
    const short al = (short)(a & 0xFFFFu), bl = (short)(b & 0xFFFFu), ah = (short)(a >> 16), bh = (short)(b >> 16);
    return (unsigned)(unsigned short)(al > bl ? al : bl) | ((unsigned)(unsigned short)(ah > bh ? ah : bh) << 16);
}
#define SB_VADD2(a, b) sb_emul_vadd2(a, b)
#define SB_VMAXS2(a, b) sb_emul_vmaxs2(a, b)
// SB_ADD2_NC on the host: the lane-wise sum, and a count of every call whose low lane would have carried into
// the high lane (the device build uses one 32-bit add there); tests/test_host_emul.py requires the count to stay 0
static long long sb_emul_carry_violations = 0;
static inline unsigned sb_emul_add2_nc(unsigned a, unsigned b)
{
    if ((a & 0xFFFFu) + (b & 0xFFFFu) > 0xFFFFu) ++sb_emul_carry_violations;
    return sb_emul_vadd2(a, b);
}
#define SB_ADD2_NC(a, b) sb_emul_add2_nc(a, b)
static inline unsigned sb_emul_brev(unsigned x)
{
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
// prmt.b32 d, x, 0, 0xBB99 (PTX ISA, generic mode: selector msb = replicate the sign of the selected byte):
// bytes 0-1 <- sign of byte 1 (bit 15), bytes 2-3 <- sign of byte 3 (bit 31)
static inline unsigned sb_emul_sign_mask2(unsigned x)
{
    return ((x >> 15) & 1u ? 0x0000FFFFu : 0u) | ((x >> 31) & 1u ? 0xFFFF0000u : 0u);
}
#define SB_BREV(x) sb_emul_brev(x)
#define SB_SIGN_MASK2(x) sb_emul_sign_mask2(x)
#else
#define SB_CONST __constant__
#define SB_LDG(p) __ldg(p)
#define SB_KERNEL(bounds) __global__ void bounds
#define SB_SHARED_STACK(name) extern __shared__ __align__(16) int name[]
#define SB_VADD2(a, b) __vadd2(a, b)
#define SB_VMAXS2(a, b) __vmaxs2(a, b)
// Packed 16-bit add where the low lane provably cannot carry: one operand is a reachable key (0 .. 4095) or a pair
// bonus (64 / 65) and the other is a stored state value (reachable, or unreachable = 0xC000 + drift <= 0xD081 as an
// unsigned lane; even a sum of two unreachable values, >= 0x8000, stays below 0xFFFF - 65).  A plain 32-bit add is
// then lane-exact, and unlike VIADD.16x2 (ALU pipe only) the compiler may issue it on the FMA pipe (IMAD.IADD).
#define SB_ADD2_NC(a, b) ((a) + (b))
#define SB_BREV(x) __brev(x)
// not __byte_perm: it keeps only 3 bits of each selector nibble (the SASS showed PRMT 0x3311, plain byte copies)
__device__ __forceinline__ unsigned sb_sign_mask2(unsigned x)
{
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(0u), "r"(0xBB99u));
    return d;
}
#define SB_SIGN_MASK2(x) sb_sign_mask2(x)
#endif

// Block shape: 768 threads per SM at 80 registers either way; after the interpreter changes of session r2i
// 128 x 6 measured 1.02x over 192 x 4 (as do 256 x 3 and 192 x 5 at 64 registers; profiles/r2_sweep.txt), and
// its finer tiles suit the short work lists of reference-rule rounds.
#ifndef SB_WALK_THREADS
#define SB_WALK_THREADS 128
#endif
#ifndef SB_WALK_MINBLOCKS
#define SB_WALK_MINBLOCKS 6
#endif
// Gene windows are kept bit-reversed (measured 1.03x in round 2, profiles/r2_sweep.txt; 0 = the round-1 form), so that the
// next leaf sits in bit 15 / bit 31 and its half-word mask is ONE PRMT (sign replication) instead of
// LOP3 + IMAD; a consumed leaf is a left shift.  Same results; verified through the host emulation.
#ifndef SB_WALK_PRMT
#define SB_WALK_PRMT 1
#endif
// The host compiler pads the leaf stream so that no op ever crosses a 16-leaf window (pad positions carry gene bit 0
// and no label; long leaf runs are split at the boundaries).  The kernels then have no general path: an op that does
// not fit the rest of the window simply opens the next one.  Without padding (0) every fifth cherry of a typical tree
// took a second, window-crossing copy of its handler: 10 % of the executed instructions came out of code that is
// otherwise cold.  Measured 0.98 - 1.01x early in round 2, 1.023x (sweep) / 1.027x (north_star step) after the
// interpreter changes of sessions r2i - r2k (profiles/r2_sweep.txt), when instruction fetch had become the second
// largest stall; the stream grows by ~12 %.  Changes the stream layout, so engine.cu (compile_tree, sb_set_tree)
// honours the same switch.
#ifndef SB_WALK_PADDED
#define SB_WALK_PADDED 1
#endif
constexpr int WALK_THREADS = SB_WALK_THREADS;
constexpr bool WALK_PADDED = SB_WALK_PADDED != 0;
constexpr int WALK_WINDOW = 16;           // leaves per gene window (two genes per 32-bit register)
// The DP stack is split.  Packed 16-bit entries (subtrees of <= 127 leaves: at most 4-5 pending at any time, pushed
// and popped all the time) live in shared memory; 32-bit entries (the few large subtrees along the spine of the tree:
// ~46 pushes per walk at 5 000 leaves, long-lived) live in per-thread local memory.  Before the split the long-lived
// 32-bit entries sat UNDER the hot 16-bit ones and doubled the shared memory a block needs (8 units at 5 000 leaves,
// 11 at 10 000, against 4 now).
constexpr int WALK_STACK32 = 8;           // 32-bit entries a thread can hold (a balanced 32 766-leaf tree needs 7)
constexpr int WALK_NEG = -(1 << 30);   // unreachable; keys stay < 2^28 (n_leaves <= 32766), so inv+inv >= INT_MIN
constexpr uint32_t SHUFFLE_DOMAIN = 0x5C0A27u;

#ifndef SB_HOST_EMUL   // device-only kernels (packing, shuffling): not part of the host emulation
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
            if (pos < n_leaves) {   // padded streams: n_leaves counts positions, pad positions have column -1
                const int col = __ldg(&walk_col[pos]);
                if (!WALK_PADDED || col >= 0) word |= ((s_rows[gl * pitch + (col >> 5)] >> (col & 31)) & 1u) << b;
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
                                                            int trait, int P, int perm_first,
                                                            uint32_t *__restrict__ labelsW,
                                                            uint8_t *__restrict__ dbg_leaf /* [P][n_leaves] or null */)
{
    extern __shared__ uint32_t s_lab[];   // [W32][T]
    const int T = (int)blockDim.x, tid = threadIdx.x;
    const int perm = blockIdx.x * T + tid;
    for (int w = 0; w < W32; ++w) s_lab[w * T + tid] = labels_leaf[w];
    if (perm >= P) return;   // no block-wide sync below
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t rnd[4] = {0, 0, 0, 0};
    for (int i = n_leaves - 1; i >= 1; --i) {
        const int s = n_leaves - 1 - i;
        // the stream belongs to the permutation's index in the whole job (perm_first + perm): a rank that walks a
        // range of the permutations draws the same labellings a single GPU would
        if ((s & 1) == 0) philox4x32_10((uint32_t)(s >> 1), (uint32_t)(perm_first + perm), (uint32_t)trait, SHUFFLE_DOMAIN, k0, k1, rnd);
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
            if (WALK_PADDED) {      // leaf_of_pos has W32p * 32 entries, -1 = pad position or past the end
                const int leaf = __ldg(&leaf_of_pos[pos]);
                if (leaf >= 0) word |= ((s_lab[(leaf >> 5) * T + tid] >> (leaf & 31)) & 1u) << b;
            } else if (pos < n_leaves) {
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

#endif  // !SB_HOST_EMUL

// ---------------------------------------------------------------- the DP
// state index: 0 AB (g=1,t=1), 1 Ab (g=1,t=0), 2 aB (g=0,t=1), 3 ab (g=0,t=0), 4 no free path
struct WalkState {
    int p[5];   // pro keys
    int a[5];   // anti keys
};

// The compiled tree program and the label bit-vectors of the current launch live in
// constant memory: every thread of a block reads the same word at the same time, so
// the compiler keeps the program counter, the leaf position, the label bits and all
// the branching on them in the uniform datapath (ULDC / UISETP / BRA.U) and the
// vector pipes only see the DP arithmetic.
// One 62 KB pool, split per tree: the uint16 program at the start, the label vectors of the current launch
// from word WalkArgs::lab_base on.  A 5 000-leaf tree (~2 700 ops, 5 KB) leaves room for 81 labellings per launch;
// a balanced 32 766-leaf tree (24 952 ops, 50 KB) still fits with 3.
constexpr int C_POOL_WORDS = 15872;
SB_CONST uint32_t c_pool[C_POOL_WORDS];
// first label word behind a program of n_ops ops (16-byte aligned)
constexpr int walk_label_base(int n_ops) { return ((n_ops + 1) / 2 + 3) / 4 * 4; }

// op = (count << 4) | type.  "16" ops work on the packed accumulators A16 / B16 (two genes
// per register, 16-bit keys), legal while the subtree has <= WALK_LIM16 leaves; "32" ops work
// on the per-gene 32-bit accumulator A32.  The host compiler (engine.cu) switches mode with
// WIDEN_A and the *W merge forms where a subtree outgrows 16 bits.
// The ops are numbered so that the interpreter finds its way with ONE single-bit test per level (a uniform
// LOP3 that sets the predicate, then the branch), most frequent first: bit 2 = B cherries (| 2: merge),
// bit 3 = A pushes / cherries (| 32: push, | 2: cherry), bit 4 = the rare 32-bit ops (a jump table on bits 0, 1
// and 5; bits 2 and 3 stay clear), else bit 5 = pops, bit 1 = single leaves.  (Compares of the whole type
// field ended up in the compiler's jump table, 12 instructions before a pop or a leaf run started; and a test of
// bit 0 compiles to LOP3 + ISETP + branch, so no frequent test uses it.)
constexpr int OP_END = 0,
              OP_MERGE_POP16 = 32,
              OP_LEAF_A16 = 2,
              OP_CHERRY_B16 = 4,        // B16 <- cherry, then `count` leaf updates
              OP_CHERRY_B16_MERGE = 6,  // as OP_CHERRY_B16, then A16 <- combine(A16, B16)
              OP_FLAG_BARE = 32,        // set on OP_CHERRY_B16_MERGE ops without further leaves (count 0)
              OP_PUSH16 = 40,           // spill A16 to the stack
              OP_CHERRY_A16 = 10,       // A16 <- cherry, then `count` leaf updates
              OP_PUSH_CHERRY_A16 = 42,  // push A16 first, then as OP_CHERRY_A16
              OP_WIDEN_A = 16, OP_LEAF_A32 = 17, OP_MERGE_A32_B16 = 18, OP_PUSH32 = 19, OP_MERGE_POP32 = 48,
              OP_MERGE_POPW = 49;
// raw (pre-fusion) steps used only by the host compiler (never stored in a program)
constexpr int RAW_LEAF_B16 = 64, RAW_MERGE_AB16 = 65;
constexpr int OP_TYPE_BITS = 6, OP_TYPE_MASK = (1 << OP_TYPE_BITS) - 1, OP_MAX_COUNT = 1023;
constexpr int PERMS_PER_ITEM_MAX = 4;      // labellings walked per block: 4, 2 or 1 (one byte of hit flags per gene)
// 16-bit keys: (pairs << 6) + x with pairs, x <= 63  ->  subtrees of at most 127 leaves.
// Unreachable = -16384 (+ drift < 4096), so valid + unreachable < 0 and unreachable + unreachable
// >= -32768 never wraps; both halves of a register follow the same rules as the 32-bit keys.
constexpr int WALK_LIM16 = 127;
constexpr int WALK_SH16 = 6;
constexpr unsigned NEG16x2 = 0xC000C000u;
constexpr unsigned K16x2 = 0x00400040u;    // one pair
constexpr unsigned K16P1x2 = 0x00410041u;  // one pair that also counts as pro (or anti)

SB_DEV int max5(const int v[5])
{
    return __vimax3_s32(__vimax3_s32(v[0], v[1], v[2]), v[3], v[4]);
}

// The DP runs one or two "passes".  K4 (DUAL = true) needs Total, Pro and Anti and keeps both key
// sets: p = (pairs, pro | pairs), a = (pairs, anti | pairs).  The permutation test (K5,
// DUAL = false) only needs Total and ONE statistic per gene -- Pro if the unpermuted Pro >= Anti,
// else Anti (methods.py:1333-1338) -- and the two passes differ only in which kind of pair adds
// the extra +1, so K5 runs a single pass with per-gene pair bonuses: half the state, half the
// arithmetic, half the stack traffic.
struct Bonus32 {     // what a supporting (AB+ab) / opposing (Ab+aB) pair adds to the p and a keys
    int ps, po, as_, ao;
};
struct Bonus16 {     // the same, packed for the two genes of a pair
    unsigned ps, po, as_, ao;
};

// ---- 32-bit (one gene) node updates
// acc <- combine(acc, leaf (g, TB)); the trait bit TB is block-uniform (the caller branches
// on it once for all genes of the thread), the gene bit is per gene
template <int TB, bool DUAL>
SB_DEV void walk_leaf(WalkState &s, bool G1, const Bonus32 &b)
{
    const int Mp = max5(s.p);
    if (TB) {  // leaf is AB (g) or aB (!g); its complement is ab (supporting pair) or Ab (opposing pair)
        const int np4 = G1 ? s.p[3] + b.ps : s.p[1] + b.po;
        s.p[0] = G1 ? Mp : s.p[0];
        s.p[2] = G1 ? s.p[2] : Mp;
        s.p[4] = np4;
    } else {   // leaf is Ab (g) or ab (!g); complement aB (opposing) or AB (supporting)
        const int np4 = G1 ? s.p[2] + b.po : s.p[0] + b.ps;
        s.p[1] = G1 ? Mp : s.p[1];
        s.p[3] = G1 ? s.p[3] : Mp;
        s.p[4] = np4;
    }
    if constexpr (DUAL) {
        const int Ma = max5(s.a);
        if (TB) {
            const int na4 = G1 ? s.a[3] + b.as_ : s.a[1] + b.ao;
            s.a[0] = G1 ? Ma : s.a[0];
            s.a[2] = G1 ? s.a[2] : Ma;
            s.a[4] = na4;
        } else {
            const int na4 = G1 ? s.a[2] + b.ao : s.a[0] + b.as_;
            s.a[1] = G1 ? Ma : s.a[1];
            s.a[3] = G1 ? s.a[3] : Ma;
            s.a[4] = na4;
        }
    }
}

SB_DEV void merge_pass(const int L[5], const int R[5], int out[5], int bsup, int bopp)
{
    const int ML = max5(L), MR = max5(R);
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = __viaddmax_s32(L[c], MR, ML + R[c]);
    // out[4] = max(L4 + R4, pp + bsup, ap + bopp, unreachable) as a chain of fused add-max steps: the clamp rides on
    // the first one and the pair bonuses on the last two (7 instructions instead of 9)
    const int nf = __viaddmax_s32(L[4], R[4], WALK_NEG);
    const int pp = __viaddmax_s32(L[0], R[3], L[3] + R[0]);
    const int ap = __viaddmax_s32(L[1], R[2], L[2] + R[1]);
    out[4] = __viaddmax_s32(ap, bopp, __viaddmax_s32(pp, bsup, nf));
}

// acc <- combine(L, acc)
template <bool DUAL>
SB_DEV void walk_merge(const WalkState &L, WalkState &acc, const Bonus32 &b)
{
    WalkState o;
    merge_pass(L.p, acc.p, o.p, b.ps, b.po);
#pragma unroll
    for (int c = 0; c < 5; ++c) acc.p[c] = o.p[c];
    if constexpr (DUAL) {
        merge_pass(L.a, acc.a, o.a, b.as_, b.ao);
#pragma unroll
        for (int c = 0; c < 5; ++c) acc.a[c] = o.a[c];
    }
}

// ---- packed 16-bit (two genes per register) node updates: the same recurrences with the
// .S16x2 DPX forms; per-gene choices become bitwise selects under a half-word mask
struct WalkState16 {
    unsigned p[5];
    unsigned a[5];
};

SB_DEV unsigned sel2(unsigned m, unsigned x, unsigned y) { return (x & m) | (y & ~m); }

SB_DEV unsigned max5_16(const unsigned v[5])
{
    return __vimax3_s16x2(__vimax3_s16x2(v[0], v[1], v[2]), v[3], v[4]);
}

// node of a trait-positive leaf (gene mask mB) and a trait-negative leaf (gene mask mb)
template <bool DUAL>
SB_DEV void walk_cherry16_mixed(WalkState16 &o, unsigned mB, unsigned mb, const Bonus16 &b)
{
    const unsigned N = NEG16x2;
    o.p[0] = N & ~mB;               // AB
    o.p[1] = N & ~mb;               // Ab
    o.p[2] = N & mB;                // aB
    o.p[3] = N & mb;                // ab
    // a pair exists where the genes differ: AB+ab (mB set) supports, aB+Ab opposes -- three LOP3 per key set
    const unsigned d = mB ^ mb;
    o.p[4] = sel2(d, sel2(mB, b.ps, b.po), N);
    if constexpr (DUAL) o.a[4] = sel2(d, sel2(mB, b.as_, b.ao), N);
}

// o[q] <- node of two leaves, for every gene pair q of the thread.  m1, m2: half-word masks of the two
// leaves' gene bits; t12 = t1 + 2*t2, the leaves' trait bits (block-uniform: one branch for all pairs)
template <int NPAIR, bool DUAL>
SB_DEV void walk_cherry16(WalkState16 *o, const unsigned (&m1)[NPAIR], const unsigned (&m2)[NPAIR],
                          unsigned t12, const Bonus16 (&b)[NPAIR])
{
    const unsigned N = NEG16x2;
    if (t12 == 3u) {                // both trait-positive: leaves in {AB, aB}
#pragma unroll
        for (int q = 0; q < NPAIR; ++q) {
            o[q].p[0] = N & ~(m1[q] | m2[q]);
            o[q].p[1] = N;
            o[q].p[2] = N & (m1[q] & m2[q]);
            o[q].p[3] = N;
            o[q].p[4] = N;
        }
    } else if (t12 == 0u) {         // both trait-negative: leaves in {Ab, ab}
#pragma unroll
        for (int q = 0; q < NPAIR; ++q) {
            o[q].p[0] = N;
            o[q].p[1] = N & ~(m1[q] | m2[q]);
            o[q].p[2] = N;
            o[q].p[3] = N & (m1[q] & m2[q]);
            o[q].p[4] = N;
        }
    } else if (t12 == 1u) {         // first leaf trait-positive, second trait-negative
#pragma unroll
        for (int q = 0; q < NPAIR; ++q) walk_cherry16_mixed<DUAL>(o[q], m1[q], m2[q], b[q]);
    } else {
#pragma unroll
        for (int q = 0; q < NPAIR; ++q) walk_cherry16_mixed<DUAL>(o[q], m2[q], m1[q], b[q]);
    }
    if constexpr (DUAL) {
#pragma unroll
        for (int q = 0; q < NPAIR; ++q) {
#pragma unroll
            for (int c = 0; c < 4; ++c) o[q].a[c] = o[q].p[c];
            if (t12 == 3u || t12 == 0u) o[q].a[4] = N;
        }
    }
}

template <int TB, bool DUAL>
SB_DEV void walk_leaf16(WalkState16 &s, unsigned m, const Bonus16 &b)
{
    const unsigned Mp = max5_16(s.p);
    if (TB) {
        const unsigned np4 = sel2(m, SB_ADD2_NC(s.p[3], b.ps), SB_ADD2_NC(s.p[1], b.po));
        s.p[0] = sel2(m, Mp, s.p[0]);
        s.p[2] = sel2(m, s.p[2], Mp);
        s.p[4] = np4;
    } else {
        const unsigned np4 = sel2(m, SB_ADD2_NC(s.p[2], b.po), SB_ADD2_NC(s.p[0], b.ps));
        s.p[1] = sel2(m, Mp, s.p[1]);
        s.p[3] = sel2(m, s.p[3], Mp);
        s.p[4] = np4;
    }
    if constexpr (DUAL) {
        const unsigned Ma = max5_16(s.a);
        if (TB) {
            const unsigned na4 = sel2(m, SB_ADD2_NC(s.a[3], b.as_), SB_ADD2_NC(s.a[1], b.ao));
            s.a[0] = sel2(m, Ma, s.a[0]);
            s.a[2] = sel2(m, s.a[2], Ma);
            s.a[4] = na4;
        } else {
            const unsigned na4 = sel2(m, SB_ADD2_NC(s.a[2], b.ao), SB_ADD2_NC(s.a[0], b.as_));
            s.a[1] = sel2(m, Ma, s.a[1]);
            s.a[3] = sel2(m, s.a[3], Ma);
            s.a[4] = na4;
        }
    }
}

SB_DEV void merge_pass16(const unsigned L[5], const unsigned R[5], unsigned out[5], unsigned bsup,
                                             unsigned bopp)
{
    const unsigned ML = max5_16(L), MR = max5_16(R);
    // out[4] as in merge_pass: clamp and pair bonuses ride on fused add-max steps.  It comes first in the source
    // because it needs the old R[0..3], which out[0..3] may then overwrite in place (callers pass acc as R).
    const unsigned nf = __viaddmax_s16x2(L[4], R[4], NEG16x2);
    const unsigned pp = __viaddmax_s16x2(L[0], R[3], SB_VADD2(L[3], R[0]));
    const unsigned ap = __viaddmax_s16x2(L[1], R[2], SB_VADD2(L[2], R[1]));
    out[4] = __viaddmax_s16x2(ap, bopp, __viaddmax_s16x2(pp, bsup, nf));
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = __viaddmax_s16x2(L[c], MR, SB_ADD2_NC(ML, R[c]));   // ML is reachable
}

// acc <- combine(L, acc)
template <bool DUAL>
SB_DEV void walk_merge16(const WalkState16 &L, WalkState16 &acc, const Bonus16 &b)
{
    WalkState16 o;
    merge_pass16(L.p, acc.p, o.p, b.ps, b.po);
#pragma unroll
    for (int c = 0; c < 5; ++c) acc.p[c] = o.p[c];
    if constexpr (DUAL) {
        merge_pass16(L.a, acc.a, o.a, b.as_, b.ao);
#pragma unroll
        for (int c = 0; c < 5; ++c) acc.a[c] = o.a[c];
    }
}

// acc <- combine(acc, cherry of two leaves with the SAME trait label), without building the cherry: such a node has
// two reachable states, both with key 0 -- "gene present" (AB or Ab) unless neither leaf carries the gene, "gene
// absent" (aB or ab) unless both do -- and no pair, so its maximum is 0 and two of acc's free-path states pass through
// unchanged: 10 instructions per gene pair instead of 5 + 20.  m1, m2: the leaves' gene masks.
template <int TB>
SB_DEV void merge_pass16_cherry_same(unsigned L[5], unsigned m1, unsigned m2, unsigned bsup, unsigned bopp)
{
    const unsigned Rx = NEG16x2 & ~(m1 | m2), Ry = NEG16x2 & (m1 & m2);
    const unsigned ML = max5_16(L);
    if (TB) {   // leaves in {AB, aB}: AB pairs with acc's ab (supporting), aB with acc's Ab (opposing)
        const unsigned x = __viaddmax_s16x2(L[3], SB_ADD2_NC(Rx, bsup), NEG16x2);
        L[4] = __viaddmax_s16x2(L[1], SB_ADD2_NC(Ry, bopp), x);
        L[0] = __viaddmax_s16x2(ML, Rx, L[0]);
        L[2] = __viaddmax_s16x2(ML, Ry, L[2]);
    } else {    // leaves in {Ab, ab}: Ab pairs with acc's aB (opposing), ab with acc's AB (supporting)
        const unsigned x = __viaddmax_s16x2(L[2], SB_ADD2_NC(Rx, bopp), NEG16x2);
        L[4] = __viaddmax_s16x2(L[0], SB_ADD2_NC(Ry, bsup), x);
        L[1] = __viaddmax_s16x2(ML, Rx, L[1]);
        L[3] = __viaddmax_s16x2(ML, Ry, L[3]);
    }
}

template <int TB, bool DUAL>
SB_DEV void walk_merge16_cherry_same(WalkState16 &acc, unsigned m1, unsigned m2, const Bonus16 &b)
{
    merge_pass16_cherry_same<TB>(acc.p, m1, m2, b.ps, b.po);
    if constexpr (DUAL) merge_pass16_cherry_same<TB>(acc.a, m1, m2, b.as_, b.ao);
}

// 16-bit key -> 32-bit key of the same (pairs, x); unreachable stays unreachable
SB_DEV int widen_key(int k16, int scale)   // scale = (1 << SH) - 64
{
    const int k32 = k16 + (k16 >> WALK_SH16) * scale;
    return k16 < 0 ? WALK_NEG : k32;
}

template <bool DUAL>
SB_DEV void walk_widen(const WalkState16 &s, WalkState &g0, WalkState &g1, int scale)
{
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        g0.p[c] = widen_key((int)(short)(s.p[c] & 0xFFFFu), scale);
        g1.p[c] = widen_key((int)s.p[c] >> 16, scale);
        if constexpr (DUAL) {
            g0.a[c] = widen_key((int)(short)(s.a[c] & 0xFFFFu), scale);
            g1.a[c] = widen_key((int)s.a[c] >> 16, scale);
        }
    }
}

// The packed accumulator A16 and the 32-bit accumulator of a pair's first gene are never live at the same time
// (WIDEN_A turns one into the other), so they share their registers: in 32-bit mode a16[s] holds the keys of gene 2s
// bit for bit and only the second gene needs registers of its own (10 fewer live registers in the interpreter loop).
template <bool DUAL>
SB_DEV WalkState as_state32(const WalkState16 &h)
{
    WalkState w;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        w.p[c] = (int)h.p[c];
        if constexpr (DUAL) w.a[c] = (int)h.a[c];
    }
    return w;
}

template <bool DUAL>
SB_DEV void put_state32(WalkState16 &h, const WalkState &w)
{
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        h.p[c] = (unsigned)w.p[c];
        if constexpr (DUAL) h.a[c] = (unsigned)w.a[c];
    }
}

struct WalkArgs {
    const uint32_t *genesT;    // [W32p][Gs]
    int64_t Gs;
    const int64_t *gene_idx;   // [S_total] or null (identity): gene row of result slot s
    const int32_t *slot_idx;   // [S] or null (identity): result slots still being walked (early-stop rounds)
    int64_t S;                 // slots walked by this launch
    int64_t S_total;           // result slots of the call (row stride of `hits`)
    int32_t W32p;
    int32_t shift;             // SH
    int32_t n_perms;           // labellings in constant memory for this launch (permute mode)
    int32_t ppi;               // labellings per block (PERMS_PER_ITEM_MAX or less when few genes are walked)
    int32_t items_per_tile;    // ceil(n_perms / ppi)
    int32_t chunk_base;        // first hit-byte row of this launch
    const int32_t *unperm;     // [S][3] (permute mode): unpermuted Total, Pro, Anti
    int32_t *pairs;            // [S][3] (pairs mode output)
    uint8_t *hits;             // [n_chunks_total][S] (permute mode output): bit r = perm ppi*chunk + r hit
                               // transposed launches: [rows][S_total] bytes, 1 = the row's gene hit under labelling s
    const int32_t *S_dev;      // if set: the number of slots to walk is read from device memory (early-stop rounds
                               // are enqueued without a host round trip; blocks past the end exit at once)
    int32_t lab_base;          // first word of the launch's label vectors in c_pool
    int32_t row_base;          // transposed launches: position (in slot_idx, or the slot itself) of constant row 0
    const int32_t *col_idx;    // if set: column of genesT that holds result slot s (genesT is then the compacted matrix of
                               // the slots still running in reference-rule mode; overrides gene_idx)
    int32_t tile_threads;      // threads per block of this launch (<= WALK_THREADS, a multiple of 32): a tile is
                               // tile_threads x NP work-list entries, chosen so that the last tile is nearly full
};

// Sweep variant: 3 = six genes per thread (12 % fewer instructions, but 96 registers leave 18 - 20 warps per SM:
// measured 0.86 - 0.91x in session r2i, 0.77x earlier in round 2; profiles/r2_sweep.txt -- not adopted)
#ifndef SB_WALK_NPAIR
#define SB_WALK_NPAIR 2
#endif
constexpr int WALK_NPAIR = SB_WALK_NPAIR;   // gene pairs per thread
constexpr int WALK_NP = 2 * WALK_NPAIR;     // genes per thread
// Words one packed push takes per thread: 5 (one pass) or 10 (both) per packed state, padded to whole 128-bit
// chunks -- an entry is stored as [chunk][thread] uint4, so a push / pop is 3 LDS.128 / STS.128 (K5: 10 words + 2 of
// padding) instead of 10 32-bit accesses.
#ifdef SB_HOST_EMUL
#define SB_HOST_DEVICE
#else
#define SB_HOST_DEVICE __host__ __device__
#endif
SB_HOST_DEVICE constexpr int walk_push_words(bool dual, int nlab) { return ((dual ? 10 : 5) * WALK_NPAIR * nlab + 3) / 4 * 4; }
// Sweep variant (18 - 20 % fewer instructions, but 128 registers leave 16 warps per SM: measured 0.99 - 1.01x in
// session r2i, 0.88 - 0.97x at 165 registers earlier in round 2, profiles/r2_sweep.txt -- not adopted): K5 walks SB_WALK_NLAB labellings of its genes in
// lockstep.  The program decode, the leaf-stream bookkeeping and the gene-bit masks are then shared by the
// labellings; only the DP arithmetic is per labelling (tools/k5_model.py counts the instructions).
#ifndef SB_WALK_NLAB
#define SB_WALK_NLAB 1
#endif
constexpr int WALK_NLAB = SB_WALK_NLAB;

// NP = 2 * NPAIR genes x NLAB labellings: simultaneous tree walks of this thread's genes under the
// labellings at c_pool[lab_off[l] ..].  gcol[k] is gene k's column of genesT (an index: four 64-bit pointers were 8 registers); genes 2q and 2q+1
// share the packed accumulators of pair q; state index s = l * NPAIR + q (32-bit: l * NP + k).  Every branch is on block-uniform data (the program and
// the label bits); per-gene data only feeds selects.  The program always ends in 32-bit mode.
// Stack entries take EW = 10 (DUAL) or 5 words: per gene pair in the packed shared-memory stack (all pairs of a
// push together, walk_push_words), per gene in the 32-bit local-memory stack (WALK_STACK32).  stk: the thread's first
// 128-bit chunk of the block's stack.
template <int NPAIR, int NLAB, bool DUAL>
SB_DEV void walk_tree(const WalkArgs &A, const uint32_t (&gcol)[2 * NPAIR], const int (&lab_off)[NLAB],
                                          int *stk, WalkState (&acc)[NLAB * 2 * NPAIR],
                                          const Bonus32 (&b32)[2 * NPAIR], const Bonus16 (&b16c)[NPAIR])
{
    constexpr int NP = 2 * NPAIR;
    constexpr int NS = NLAB * NPAIR;   // packed states
    constexpr int T = WALK_THREADS;
    constexpr int EW = DUAL ? 10 : 5;
    const int K = 1 << A.shift;
    const int scale = K - (1 << WALK_SH16);
    WalkState16 a16[NS], b16[NS];      // a16[s]: packed A of pair s, or (32-bit mode) the keys of gene 2s
    WalkState hi[NS];                  // 32-bit mode: the keys of gene 2s + 1
    constexpr int PW4 = walk_push_words(DUAL, NLAB) / 4;   // 128-bit chunks per push
    uint4 *top = reinterpret_cast<uint4 *>(stk);           // the thread's next free chunk (chunks are T apart)
    int pc = 0;                        // a BYTE offset into c_pool
    int sp32 = 0;                      // 32-bit entries held in stk32
    int stk32[WALK_STACK32 * EW * NLAB * NP];   // local memory: touched by PUSH32 / MERGE_POP32 only
    int room = 0, win = 0;             // leaves left in the current 16-leaf window; windows opened so far
    // gx[q]: the current 16-leaf window of pair q's two genes (gene 2q in bits 0..15, gene 2q+1 in
    // bits 16..31, consumed from bit 0 / bit 16); gy[q]: the following 16 leaves; gnext: prefetch
    uint32_t gx[NPAIR], gy[NPAIR], gnext[NP], lw[NLAB];
#pragma unroll
    for (int l = 0; l < NLAB; ++l) lw[l] = 0;
#pragma unroll
    for (int k = 0; k < NP; ++k) gnext[k] = SB_LDG(A.genesT + gcol[k]);
#pragma unroll
    for (int q = 0; q < NPAIR; ++q) { gx[q] = 0; gy[q] = 0; }
    const int W32p = A.W32p;
    const int64_t Gs = A.Gs;
    // The leaf stream.  lw holds the label bits from the next leaf on (bit 0 = the next leaf) and gx[q] the
    // gene bits of pair q up to the end of its 16-leaf window (bit 0 / bit 16 = the next leaf); `room`
    // leaves of the window are left.  An op whose leaves all lie in the window (the usual case: one
    // block-uniform test per op) consumes them without any per-leaf test; an op that crosses into the
    // next window takes the general path, which opens a window wherever room == 0.
#define SB_OPEN_WINDOW()                                                                       \
    do {                                                                                       \
        if ((win & 1) == 0) {                                                                  \
            const int w_ = win >> 1;                                                           \
            _Pragma("unroll") for (int l_ = 0; l_ < NLAB; ++l_) lw[l_] = c_pool[lab_off[l_] + w_]; \
            _Pragma("unroll") for (int q_ = 0; q_ < NPAIR; ++q_) {                             \
                const uint32_t a_ = gnext[2 * q_], b_ = gnext[2 * q_ + 1];                     \
                gx[q_] = SB_WINDOW_X(a_, b_);                                                  \
                gy[q_] = SB_WINDOW_Y(a_, b_);                                                  \
            }                                                                                  \
            if (w_ + 1 < W32p) {                                                               \
                _Pragma("unroll") for (int k_ = 0; k_ < NP; ++k_)                              \
                    gnext[k_] = SB_LDG(A.genesT + (int64_t)(w_ + 1) * Gs + gcol[k_]);        \
            }                                                                                  \
        } else {                                                                               \
            _Pragma("unroll") for (int q_ = 0; q_ < NPAIR; ++q_) gx[q_] = gy[q_];              \
            if (WALK_PADDED) {   /* skipped pad positions: the label word is not where the shifts left it */ \
                _Pragma("unroll") for (int l_ = 0; l_ < NLAB; ++l_)                            \
                    lw[l_] = c_pool[lab_off[l_] + (win >> 1)] >> 16;                           \
            }                                                                                  \
        }                                                                                      \
        ++win;                                                                                 \
        room = WALK_WINDOW;                                                                    \
    } while (0)
    // SB_PAIR_MASK: half-word mask of pair q's gene bits at stream offset o (0 = the next leaf), 0xFFFF per
    // half whose gene is present; SB_CONSUME: drop n leaves from pair q's window
#if SB_WALK_PRMT
#define SB_WINDOW_X(a_, b_) ((SB_BREV(a_) >> 16) | (SB_BREV(b_) & 0xFFFF0000u))
#define SB_WINDOW_Y(a_, b_) ((SB_BREV(a_) & 0xFFFFu) | (SB_BREV(b_) << 16))
#define SB_PAIR_MASK(q, o) SB_SIGN_MASK2(gx[q] << (o))
#define SB_GENE_BIT(k) (((gx[(k) >> 1] >> (((k) & 1) * 16 + 15)) & 1u) != 0)
#define SB_CONSUME(q, n) gx[q] <<= (n)
#else
#define SB_WINDOW_X(a_, b_) (((a_) & 0xFFFFu) | ((b_) << 16))
#define SB_WINDOW_Y(a_, b_) (((a_) >> 16) | ((b_) & 0xFFFF0000u))
#define SB_PAIR_MASK(q, o) (((gx[q] >> (o)) & 0x00010001u) * 0xFFFFu)
#define SB_GENE_BIT(k) (((gx[(k) >> 1] >> (((k) & 1) * 16)) & 1u) != 0)
#define SB_CONSUME(q, n) gx[q] >>= (n)
#endif
    // one leaf update of the packed / the 32-bit accumulators; `room` is the caller's business
#define SB_LEAF_STEP16(ACC)                                                                    \
    do {                                                                                       \
        unsigned mk_[NPAIR];                                                                   \
        _Pragma("unroll") for (int q = 0; q < NPAIR; ++q) mk_[q] = SB_PAIR_MASK(q, 0);         \
        _Pragma("unroll") for (int l = 0; l < NLAB; ++l) {                                     \
            const uint32_t t_ = lw[l] & 1u;                                                    \
            lw[l] >>= 1;                                                                       \
            if (t_) {                                                                          \
                _Pragma("unroll") for (int q = 0; q < NPAIR; ++q)                              \
                    walk_leaf16<1, DUAL>(ACC[l * NPAIR + q], mk_[q], b16c[q]);                 \
            } else {                                                                           \
                _Pragma("unroll") for (int q = 0; q < NPAIR; ++q)                              \
                    walk_leaf16<0, DUAL>(ACC[l * NPAIR + q], mk_[q], b16c[q]);                 \
            }                                                                                  \
        }                                                                                      \
        _Pragma("unroll") for (int q = 0; q < NPAIR; ++q) SB_CONSUME(q, 1);                    \
    } while (0)
#define SB_LEAF_STEP32()                                                                       \
    do {                                                                                       \
        _Pragma("unroll") for (int l = 0; l < NLAB; ++l) {                                     \
            const uint32_t t_ = lw[l] & 1u;                                                    \
            lw[l] >>= 1;                                                                       \
            _Pragma("unroll") for (int q = 0; q < NPAIR; ++q) {                                \
                WalkState g0_ = as_state32<DUAL>(a16[l * NPAIR + q]);                          \
                if (t_) {                                                                      \
                    walk_leaf<1, DUAL>(g0_, SB_GENE_BIT(2 * q), b32[2 * q]);                   \
                    walk_leaf<1, DUAL>(hi[l * NPAIR + q], SB_GENE_BIT(2 * q + 1), b32[2 * q + 1]); \
                } else {                                                                       \
                    walk_leaf<0, DUAL>(g0_, SB_GENE_BIT(2 * q), b32[2 * q]);                   \
                    walk_leaf<0, DUAL>(hi[l * NPAIR + q], SB_GENE_BIT(2 * q + 1), b32[2 * q + 1]); \
                }                                                                              \
                put_state32<DUAL>(a16[l * NPAIR + q], g0_);                                    \
            }                                                                                  \
        }                                                                                      \
        _Pragma("unroll") for (int q = 0; q < NPAIR; ++q) SB_CONSUME(q, 1);                    \
    } while (0)
    // `cnt` leaf updates
#define SB_LEAF_RUN(STEP)                                                                      \
    do {                                                                                       \
        if (WALK_PADDED) {   /* the compiler never lets a run cross a window: what is left is padding */ \
            if (cnt > room) SB_OPEN_WINDOW();                                                  \
            room -= cnt;                                                                       \
            _Pragma("unroll 1") for (int i = 0; i < cnt; ++i) { STEP; }                        \
        } else if (__builtin_expect(cnt <= room, 1)) {                                         \
            room -= cnt;                                                                       \
            _Pragma("unroll 1") for (int i = 0; i < cnt; ++i) { STEP; }                        \
        } else {                                                                               \
            _Pragma("unroll 1") for (int i = 0; i < cnt; ++i) {                                \
                if (room == 0) SB_OPEN_WINDOW();                                               \
                STEP;                                                                          \
                --room;                                                                        \
            }                                                                                  \
        }                                                                                      \
    } while (0)
    // ACC <- node of the next two leaves (t12: label of the first in bit 0, of the second in bit 1), then
    // `cnt` leaf updates of ACC
    // FUSE (the B cherries of single-labelling kernels): a bare cherry that is merged into A at once (the most
    // frequent op of a typical tree, flagged by the host compiler) skips B when its leaves carry the same label
#define SB_CHERRY_RUN16(ACC, FUSE)                                                             \
    {                                                                                          \
        uint32_t t12[NLAB];                                                                    \
        unsigned m1[NPAIR], m2[NPAIR];                                                         \
        if (WALK_PADDED && cnt + 2 > room) SB_OPEN_WINDOW();   /* the rest of the window is padding */ \
        if (WALK_PADDED || __builtin_expect(cnt + 2 <= room, 1)) {                             \
            room -= cnt + 2;                                                                   \
            _Pragma("unroll") for (int l = 0; l < NLAB; ++l) {                                 \
                t12[l] = lw[l] & 3u;                                                           \
                lw[l] >>= 2;                                                                   \
            }                                                                                  \
            _Pragma("unroll") for (int q = 0; q < NPAIR; ++q) {                                \
                m1[q] = SB_PAIR_MASK(q, 0);                                                    \
                m2[q] = SB_PAIR_MASK(q, 1);                                                    \
                SB_CONSUME(q, 2);                                                              \
            }                                                                                  \
            if (FUSE && (op & (uint32_t)OP_FLAG_BARE)) {                                       \
                if (t12[0] == 3u) {                                                            \
                    _Pragma("unroll") for (int q = 0; q < NPAIR; ++q)                          \
                        walk_merge16_cherry_same<1, DUAL>(a16[q], m1[q], m2[q], b16c[q]);      \
                    continue;                                                                  \
                }                                                                              \
                if (t12[0] == 0u) {                                                            \
                    _Pragma("unroll") for (int q = 0; q < NPAIR; ++q)                          \
                        walk_merge16_cherry_same<0, DUAL>(a16[q], m1[q], m2[q], b16c[q]);      \
                    continue;                                                                  \
                }                                                                              \
            }                                                                                  \
            _Pragma("unroll") for (int l = 0; l < NLAB; ++l)                                   \
                walk_cherry16<NPAIR, DUAL>(ACC + l * NPAIR, m1, m2, t12[l], b16c);             \
            if (cnt) {                                                                         \
                int i = cnt;                                                                   \
                _Pragma("unroll 1") do { SB_LEAF_STEP16(ACC); } while (--i);                   \
            }                                                                                  \
        } else {                                                                               \
            if (room == 0) SB_OPEN_WINDOW();                                                   \
            _Pragma("unroll") for (int l = 0; l < NLAB; ++l) {                                 \
                t12[l] = lw[l] & 1u;                                                           \
                lw[l] >>= 1;                                                                   \
            }                                                                                  \
            _Pragma("unroll") for (int q = 0; q < NPAIR; ++q) {                                \
                m1[q] = SB_PAIR_MASK(q, 0);                                                    \
                SB_CONSUME(q, 1);                                                              \
            }                                                                                  \
            if (--room == 0) SB_OPEN_WINDOW();                                                 \
            _Pragma("unroll") for (int l = 0; l < NLAB; ++l) {                                 \
                t12[l] |= (lw[l] & 1u) << 1;                                                   \
                lw[l] >>= 1;                                                                   \
            }                                                                                  \
            _Pragma("unroll") for (int q = 0; q < NPAIR; ++q) {                                \
                m2[q] = SB_PAIR_MASK(q, 0);                                                    \
                SB_CONSUME(q, 1);                                                              \
            }                                                                                  \
            --room;                                                                            \
            _Pragma("unroll") for (int l = 0; l < NLAB; ++l)                                   \
                walk_cherry16<NPAIR, DUAL>(ACC + l * NPAIR, m1, m2, t12[l], b16c);             \
            _Pragma("unroll 1") for (int i = 0; i < cnt; ++i) {                                \
                if (room == 0) SB_OPEN_WINDOW();                                               \
                SB_LEAF_STEP16(ACC);                                                           \
                --room;                                                                        \
            }                                                                                  \
        }                                                                                      \
    }
    // pop one entry into Ls[NS]; word q * EW + c (+ 5 for the second pass) of a push belongs to state q
#define SB_POP16_ALL(Ls)                                                                       \
    do {                                                                                       \
        unsigned w_[PW4 * 4];                                                                  \
        top -= PW4 * T;                                                                        \
        _Pragma("unroll") for (int j = 0; j < PW4; ++j) {                                      \
            const uint4 v_ = top[j * T];                                                       \
            w_[4 * j] = v_.x; w_[4 * j + 1] = v_.y; w_[4 * j + 2] = v_.z; w_[4 * j + 3] = v_.w; \
        }                                                                                      \
        _Pragma("unroll") for (int q = 0; q < NS; ++q) {                                       \
            _Pragma("unroll") for (int c = 0; c < 5; ++c) {                                    \
                Ls[q].p[c] = w_[q * EW + c];                                                   \
                if constexpr (DUAL) Ls[q].a[c] = w_[q * EW + 5 + c];                           \
            }                                                                                  \
        }                                                                                      \
    } while (0)
    for (;;) {
        const uint32_t op = *reinterpret_cast<const uint16_t *>(reinterpret_cast<const char *>(c_pool) + pc);
        pc += 2;
        const int cnt = op >> OP_TYPE_BITS;
        if (op & 4u) {
            SB_CHERRY_RUN16(b16, NLAB == 1);
            if (op & 2u) {
#pragma unroll
                for (int s = 0; s < NS; ++s) walk_merge16<DUAL>(b16[s], a16[s], b16c[s % NPAIR]);
            }
            continue;
        }
        if (op & 8u) {
            if (op & 32u) {
                unsigned w[PW4 * 4];
#pragma unroll
                for (int k = EW * NS; k < PW4 * 4; ++k) w[k] = 0u;
#pragma unroll
                for (int q = 0; q < NS; ++q) {
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        w[q * EW + c] = a16[q].p[c];
                        if constexpr (DUAL) w[q * EW + 5 + c] = a16[q].a[c];
                    }
                }
#pragma unroll
                for (int j = 0; j < PW4; ++j) top[j * T] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                top += PW4 * T;
            }
            if (op & 2u) SB_CHERRY_RUN16(a16, false);
            continue;
        }
        if ((op & 16u) == 0) {
            if (op & 32u) {         // OP_MERGE_POP16 (always a single pop)
                WalkState16 L[NS];
                SB_POP16_ALL(L);
#pragma unroll
                for (int q = 0; q < NS; ++q) walk_merge16<DUAL>(L[q], a16[q], b16c[q % NPAIR]);
                continue;
            }
            if (op & 2u) {          // OP_LEAF_A16
                SB_LEAF_RUN(SB_LEAF_STEP16(a16));
                continue;
            }
            break;                  // OP_END
        }
        switch (op & OP_TYPE_MASK) {
        case OP_WIDEN_A:
#pragma unroll
            for (int q = 0; q < NS; ++q) {
                WalkState g0;
                walk_widen<DUAL>(a16[q], g0, hi[q], scale);
                put_state32<DUAL>(a16[q], g0);
            }
            break;
        case OP_LEAF_A32:
            SB_LEAF_RUN(SB_LEAF_STEP32());
            break;
        case OP_MERGE_A32_B16:
#pragma unroll
            for (int q = 0; q < NS; ++q) {
                WalkState r0, r1, g0 = as_state32<DUAL>(a16[q]);
                walk_widen<DUAL>(b16[q], r0, r1, scale);
                walk_merge<DUAL>(r0, g0, b32[(2 * q) % NP]);
                walk_merge<DUAL>(r1, hi[q], b32[(2 * q + 1) % NP]);
                put_state32<DUAL>(a16[q], g0);
            }
            break;
        case OP_PUSH32: {
            int *s = stk32 + sp32 * (EW * NLAB * NP);
#pragma unroll
            for (int q = 0; q < NS; ++q) {
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    s[2 * q * EW + c] = (int)a16[q].p[c];
                    s[(2 * q + 1) * EW + c] = hi[q].p[c];
                    if constexpr (DUAL) {
                        s[2 * q * EW + 5 + c] = (int)a16[q].a[c];
                        s[(2 * q + 1) * EW + 5 + c] = hi[q].a[c];
                    }
                }
            }
            ++sp32;
            break;
        }
        case OP_MERGE_POP32:
#pragma unroll 1
            for (int i = 0; i < cnt; ++i) {
                --sp32;
                const int *s = stk32 + sp32 * (EW * NLAB * NP);
#pragma unroll
                for (int q = 0; q < NS; ++q) {
                    WalkState L0, L1, g0 = as_state32<DUAL>(a16[q]);
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        L0.p[c] = s[2 * q * EW + c];
                        L1.p[c] = s[(2 * q + 1) * EW + c];
                        if constexpr (DUAL) {
                            L0.a[c] = s[2 * q * EW + 5 + c];
                            L1.a[c] = s[(2 * q + 1) * EW + 5 + c];
                        }
                    }
                    walk_merge<DUAL>(L0, g0, b32[(2 * q) % NP]);
                    walk_merge<DUAL>(L1, hi[q], b32[(2 * q + 1) % NP]);
                    put_state32<DUAL>(a16[q], g0);
                }
            }
            break;
        case OP_MERGE_POPW:
#pragma unroll 1
            for (int i = 0; i < cnt; ++i) {
                WalkState16 Ls[NS];
                SB_POP16_ALL(Ls);
#pragma unroll
                for (int q = 0; q < NS; ++q) {
                    const WalkState16 &L = Ls[q];
                    WalkState r0, r1, g0 = as_state32<DUAL>(a16[q]);
                    walk_widen<DUAL>(L, r0, r1, scale);
                    walk_merge<DUAL>(r0, g0, b32[(2 * q) % NP]);
                    walk_merge<DUAL>(r1, hi[q], b32[(2 * q + 1) % NP]);
                    put_state32<DUAL>(a16[q], g0);
                }
            }
            break;
        default:
            break;
        }
    }
#pragma unroll
    for (int q = 0; q < NS; ++q) {     // the program always ends in 32-bit mode
        acc[2 * q] = as_state32<DUAL>(a16[q]);
        acc[2 * q + 1] = hi[q];
    }
#undef SB_OPEN_WINDOW
#undef SB_WINDOW_X
#undef SB_WINDOW_Y
#undef SB_PAIR_MASK
#undef SB_GENE_BIT
#undef SB_CONSUME
#undef SB_LEAF_STEP16
#undef SB_LEAF_STEP32
#undef SB_LEAF_RUN
#undef SB_CHERRY_RUN16
#undef SB_POP16_ALL
}

// gene slots of this thread: (tile * NP + k) * tile_threads + tid  (coalesced per k)
template <int NP>
SB_DEV void walk_slots(const WalkArgs &A, const int32_t *list, int tile, int32_t (&s_idx)[NP], bool (&active)[NP],
                                           uint32_t (&gcol)[NP])
{
    const int64_t S = A.S_dev ? (int64_t)*A.S_dev : A.S;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const int64_t li = ((int64_t)tile * NP + k) * A.tile_threads + threadIdx.x;   // position in the work list
        active[k] = li < S;
        const int64_t lc = active[k] ? li : (S - 1);   // idle lanes redo the last entry (no divergence)
        s_idx[k] = list ? list[lc] : (int32_t)lc;         // result slots are counted in 32 bits throughout
        const int64_t gene = A.col_idx ? (int64_t)A.col_idx[s_idx[k]] : (A.gene_idx ? A.gene_idx[s_idx[k]] : (int64_t)s_idx[k]);
        gcol[k] = (uint32_t)gene;   // columns are counted in 32 bits (sb_set_genes checks)
    }
}

// K4: one labelling (the first row behind lab_base), both passes; writes pairs[S][3] = Total, Pro, Anti.
// The root takes three independent maxima (classes.py:246-249).
SB_KERNEL(__launch_bounds__(WALK_THREADS)) walk_pairs_kernel(const WalkArgs A)
{
    SB_SHARED_STACK(smem_stack);
    int *stk = smem_stack + 4 * threadIdx.x;
    constexpr int NP = WALK_NP;
    int32_t s_idx[NP]; bool active[NP]; uint32_t gcol[NP];
    walk_slots<NP>(A, A.slot_idx, blockIdx.x, s_idx, active, gcol);
    const int K = 1 << A.shift;
    Bonus32 b32[NP];
    Bonus16 b16[WALK_NPAIR];
#pragma unroll
    for (int k = 0; k < NP; ++k) b32[k] = Bonus32{K + 1, K, K, K + 1};
#pragma unroll
    for (int q = 0; q < WALK_NPAIR; ++q) b16[q] = Bonus16{K16P1x2, K16x2, K16x2, K16P1x2};
    WalkState acc[NP];
    const int lab_off[1] = {A.lab_base};
    walk_tree<WALK_NPAIR, 1, true>(A, gcol, lab_off, stk, acc, b32, b16);
    const int mask = K - 1;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const int total = max5(acc[k].p) >> A.shift;
        int pro = -1, anti = -1;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            if (acc[k].p[c] >= 0) pro = max(pro, acc[k].p[c] & mask);
            if (acc[k].a[c] >= 0) anti = max(anti, acc[k].a[c] & mask);
        }
        if (active[k]) {
            A.pairs[(int64_t)s_idx[k] * 3 + 0] = total;
            A.pairs[(int64_t)s_idx[k] * 3 + 1] = pro;
            A.pairs[(int64_t)s_idx[k] * 3 + 2] = anti;
        }
    }
}

// K5: grid = (gene tiles, chunks of PERMS_PER_ITEM labellings).  A thread walks its NP genes
// under each labelling of the block's chunk -- a single pass keyed on each gene's own statistic --
// and writes one byte of hit flags per gene.  Blocks are small work items, so the tail of a
// launch is short, and concurrently running blocks read the same few label vectors from the
// constant cache.
//
// TRANSPOSED launches serve the usual command-line case after filtering -- a few dozen genes, thousands of
// permutations (methods.py:1022-1024, :1295-1310) -- where threads = genes would leave most of every block idle.
// The DP is symmetric in the gene bit and the trait bit of a leaf (swapping them exchanges the states Ab and aB,
// and AB+ab / Ab+aB pairs keep their kind, classes.py:459-572), so the same walk runs with the roles exchanged:
// a thread's "gene" columns are labellings (genesT = the label vectors transposed, [W32p][Ps]), the constant rows
// are the walk-order bits of up to ppi genes, and the tested side / unpermuted counts belong to the row.
template <bool TRANSPOSED>
SB_KERNEL(__launch_bounds__(WALK_THREADS, SB_WALK_MINBLOCKS)) walk_permute_kernel(const WalkArgs A)
{
    SB_SHARED_STACK(smem_stack);
    int *stk = smem_stack + 4 * threadIdx.x;
    constexpr int NP = WALK_NP;
    constexpr int NLAB = TRANSPOSED ? 1 : WALK_NLAB;
    const int chunk = blockIdx.y;
    const int perm0 = chunk * A.ppi;
    const int rows = min(A.ppi, A.n_perms - perm0);
    if (A.S_dev && (int64_t)blockIdx.x * NP * A.tile_threads >= (int64_t)*A.S_dev) return;   // grid sized for an upper bound
    int32_t s_idx[NP]; bool active[NP]; uint32_t gcol[NP];
    walk_slots<NP>(A, TRANSPOSED ? nullptr : A.slot_idx, blockIdx.x, s_idx, active, gcol);   // transposed: slot_idx lists the rows
    const int K = 1 << A.shift;
    const int mask = K - 1;
    int u_total[NP], u_stat[NP]; bool use_pro[NP]; uint32_t hits[NP];      // unpermuted counts (<= 16 383)
    Bonus32 b32[NP];
    Bonus16 b16[WALK_NPAIR];
    // tested side and pair bonuses of "gene" k from the unpermuted counts at un: the statistic a gene is tested on
    // counts +1 for its own kind of pair (methods.py:1333-1336)
    auto side = [&](int k, const int32_t *un) {
        u_total[k] = un[0];
        use_pro[k] = un[1] >= un[2];
        u_stat[k] = use_pro[k] ? un[1] : un[2];
        b32[k] = Bonus32{use_pro[k] ? K + 1 : K, use_pro[k] ? K : K + 1, 0, 0};
    };
    auto pack_bonus = [&]() {
#pragma unroll
        for (int q = 0; q < WALK_NPAIR; ++q) {
            const unsigned s0 = use_pro[2 * q] ? 65u : 64u, s1 = use_pro[2 * q + 1] ? 65u : 64u;
            b16[q] = Bonus16{s0 | (s1 << 16), (129u - s0) | ((129u - s1) << 16), 0u, 0u};
        }
    };
#pragma unroll
    for (int k = 0; k < NP; ++k) hits[k] = 0;
    if constexpr (!TRANSPOSED) {
#pragma unroll
        for (int k = 0; k < NP; ++k) side(k, A.unperm + (int64_t)s_idx[k] * 3);
        pack_bonus();
    }
    for (int r = 0; r < rows; r += NLAB) {
        int64_t row_slot = 0;
        if constexpr (TRANSPOSED) {   // the row's gene decides the tested side for every labelling of the thread
            const int64_t e = (int64_t)A.row_base + perm0 + r;
            row_slot = A.slot_idx ? (int64_t)A.slot_idx[e] : e;
#pragma unroll
            for (int k = 0; k < NP; ++k) side(k, A.unperm + row_slot * 3);
            pack_bonus();
        }
        WalkState acc[NLAB * NP];
        int lab_off[NLAB];
#pragma unroll
        for (int l = 0; l < NLAB; ++l)   // an odd tail walks its last labelling twice
            lab_off[l] = A.lab_base + (perm0 + min(r + l, rows - 1)) * A.W32p;
        walk_tree<WALK_NPAIR, NLAB, false>(A, gcol, lab_off, stk, acc, b32, b16);
#pragma unroll
        for (int l = 0; l < NLAB; ++l) {
            if (r + l >= rows) break;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                // root: Total and the statistic are independent maxima over the five states (classes.py:246-249)
                const int total = max5(acc[l * NP + k].p) >> A.shift;
                int stat = -1;
#pragma unroll
                for (int c = 0; c < 5; ++c)
                    if (acc[l * NP + k].p[c] >= 0) stat = max(stat, acc[l * NP + k].p[c] & mask);
                const bool hit = (long long)stat * u_total[k] >= (long long)u_stat[k] * total;   // methods.py:1353-1355
                if constexpr (TRANSPOSED) {
                    if (active[k]) A.hits[row_slot * A.S_total + (int64_t)s_idx[k]] = (uint8_t)hit;
                } else {
                    if (hit) hits[k] |= (1u << (r + l));
                }
            }
        }
    }
    if constexpr (!TRANSPOSED) {
#pragma unroll
        for (int k = 0; k < NP; ++k)
            if (active[k]) A.hits[(int64_t)(A.chunk_base + chunk) * A.S_total + (int64_t)s_idx[k]] = (uint8_t)hits[k];
    }
}

#ifndef SB_HOST_EMUL   // device-only kernels below
// ---------------------------------------------------------------- hit-sequence reduction
// Permute's bookkeeping (methods.py:1348-1365), slice by slice.  `hits` holds the flags of the slice of
// permutations [base, base + n) just walked for the slots in list_in (row c, bit b = permutation base + ppi*c + b).
// Exhaustive mode adds the hits up; reference-rule mode applies the sequential stop rule (:1360-1363), and slots
// that neither stopped nor reached P are appended to list_out for the next slice.  n_in_dev: the length of
// list_in in device memory (rounds are enqueued without a host round trip), else n_in; `walks` (optional)
// accumulates the walks such a round really did, for sb_stats.
__global__ void __launch_bounds__(256) accumulate_hits_kernel(const uint8_t *__restrict__ hits, int64_t S_total,
                                                              const int32_t *__restrict__ list_in, int32_t n_in,
                                                              const int32_t *__restrict__ n_in_dev, int base, int n,
                                                              int P, int ppi, int early_stop,
                                                              const int32_t *__restrict__ rmin, int32_t *__restrict__ r_arr,
                                                              int32_t *__restrict__ n_done, int32_t *__restrict__ list_out,
                                                              int32_t *__restrict__ counter,
                                                              unsigned long long *__restrict__ walks)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_live = n_in_dev ? *n_in_dev : n_in;
    if (e == 0 && walks) atomicAdd(walks, (unsigned long long)n_live * (unsigned long long)n);
    const bool live = e < n_live;
    const int slot = live ? (list_in ? list_in[e] : e) : 0;
    if (!early_stop) {
        if (!live) return;
        int r = (base == 0) ? 0 : r_arr[slot];
        for (int c = 0; c * ppi < n; ++c)
            r += __popc((uint32_t)hits[(int64_t)c * S_total + slot] & ((1u << min(ppi, n - c * ppi)) - 1u));
        r_arr[slot] = r;
        if (base + n >= P) n_done[slot] = P;
        return;
    }
    bool survives = false;
    if (live) {
        int r = (base == 0) ? 0 : r_arr[slot];
        bool stopped = false;
        for (int k = 0; k < n; ++k) {
            const int i = base + k;
            r += (hits[(int64_t)(k / ppi) * S_total + slot] >> (k % ppi)) & 1u;
            if (i >= 30 && r >= rmin[i]) {
                n_done[slot] = i + 1;
                stopped = true;
                break;
            }
        }
        r_arr[slot] = r;
        if (!stopped) {
            if (base + n >= P) n_done[slot] = P;
            else survives = true;
        }
    }
    // warp-aggregated append: one atomic per warp, and the slots of a warp keep their order, so the next round's
    // threads read (nearly) consecutive columns
    const unsigned m = __ballot_sync(0xffffffffu, survives);
    if (m) {
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        int at = 0;
        if (lane == leader) at = atomicAdd(counter, __popc(m));
        at = __shfl_sync(0xffffffffu, at, leader);
        if (survives) list_out[at + __popc(m & ((1u << lane) - 1u))] = slot;
    }
}

// ---------------------------------------------------------------- transposed launches (few genes x many permutations)
// labelsW [P][W32p] -> labelsT [W32p][Ps]: the label vectors become the per-thread columns of a transposed K5 launch
__global__ void __launch_bounds__(256) transpose_labels_kernel(const uint32_t *__restrict__ labelsW, int P, int W32p,
                                                               int64_t Ps, uint32_t *__restrict__ labelsT)
{
    __shared__ uint32_t tile[32][33];
    const int p0 = blockIdx.x * 32, w0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int p = p0 + j, w = w0 + tx;
        tile[j][tx] = (p < P && w < W32p) ? labelsW[(int64_t)p * W32p + w] : 0u;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int w = w0 + j, p = p0 + tx;
        if (w < W32p && p < Ps) labelsT[(int64_t)w * Ps + p] = tile[tx][j];
    }
}

// rowsW [n][W32p]: the walk-order bits of the genes in result slots list[e] (or e), gathered out of genesT [W32p][Gs]
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint32_t *__restrict__ genesT, int64_t Gs, int W32p,
                                                          const int64_t *__restrict__ gene_idx,
                                                          const int32_t *__restrict__ list, int n,
                                                          uint32_t *__restrict__ rowsW)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n * W32p) return;
    const int e = (int)(i / W32p), w = (int)(i - (int64_t)e * W32p);
    const int64_t slot = list ? list[e] : e;
    const int64_t gene = gene_idx ? gene_idx[slot] : slot;
    rowsW[i] = genesT[(int64_t)w * Gs + gene];
}

// genesC [W32p][n_pad]: the columns of genesT that belong to the result slots list[e], e < *n (device count), packed
// side by side, and col_of_slot[list[e]] = e.  Reference-rule mode walks the few slots that survive the first
// rounds hundreds of times: out of the full matrix every thread would fetch its own 32-byte sector per word.
__global__ void __launch_bounds__(256) compact_columns_kernel(const uint32_t *__restrict__ genesT, int64_t Gs, int W32p,
                                                              const int64_t *__restrict__ gene_idx,
                                                              const int32_t *__restrict__ list, int n, int64_t n_pad,
                                                              uint32_t *__restrict__ genesC,
                                                              int32_t *__restrict__ col_of_slot)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int slot = list[e];
    const int64_t gene = gene_idx ? gene_idx[slot] : slot;
    if (blockIdx.y == 0) col_of_slot[slot] = e;
    for (int w = blockIdx.y; w < W32p; w += gridDim.y) genesC[(int64_t)w * n_pad + e] = genesT[(int64_t)w * Gs + gene];
}

// Permute's bookkeeping (methods.py:1348-1365) on hit rows [slot][Pp] (one byte per labelling), one warp per gene.
// n_avail labellings have been walked so far: a gene that neither stopped nor reached P goes to list_out.
__global__ void __launch_bounds__(256) reduce_hit_rows_kernel(const uint8_t *__restrict__ hits, int64_t Pp,
                                                              const int32_t *__restrict__ list, int n, int n_avail,
                                                              int P, int early_stop, const int32_t *__restrict__ rmin,
                                                              int32_t *__restrict__ r_out, int32_t *__restrict__ n_done,
                                                              int32_t *__restrict__ list_out, int32_t *__restrict__ counter)
{
    const int e = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (e >= n) return;
    const int slot = list ? list[e] : e;
    const uint8_t *row = hits + (int64_t)slot * Pp;
    int r = 0, done = -1;
    for (int base = 0; base < n_avail; base += 32) {
        const int i = base + lane;
        const bool bit = i < n_avail && row[i] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, bit);
        if (early_stop) {
            const int cum = r + __popc(m & (0xffffffffu >> (31 - lane)));
            const bool stop = i < n_avail && i >= 30 && cum >= rmin[i];     // methods.py:1360-1363
            const unsigned sm = __ballot_sync(0xffffffffu, stop);
            if (sm) {
                const int first = __ffs(sm) - 1;
                r += __popc(m & (0xffffffffu >> (31 - first)));
                done = base + first + 1;
                break;
            }
        }
        r += __popc(m);
    }
    if (lane) return;
    if (done < 0 && n_avail < P) {
        list_out[atomicAdd(counter, 1)] = slot;
        return;
    }
    r_out[slot] = r;
    n_done[slot] = done < 0 ? P : done;
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


// pipe-rate probes for kernel design decisions (sb_debug_pipe_rates)
template <int MODE>
__global__ void __launch_bounds__(256) pipe_rate_kernel(int *out, int iters, int seed)
{
    unsigned x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = seed + threadIdx.x * (k + 1);
    const unsigned a = seed | 0x00010001u, b = 0xC000C000u + seed;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 0) x[k] = __viaddmax_s32((int)x[k], (int)a, (int)(b + k));
                if (MODE == 1) x[k] = __viaddmax_s16x2(x[k], a, b + k);
                if (MODE == 2) x[k] = __vimax3_s16x2(x[k], a + k, b);
                if (MODE == 3) x[k] = SB_VADD2(x[k], a + k);
                if (MODE == 4) x[k] = (x[k] & a) | (b & ~a) ^ (x[k] >> 1);          // LOP3 + SHF
                if (MODE == 5) x[k] = __vimax3_s32((int)x[k], (int)(a + k), (int)b);
                if (MODE == 6) x[k] = (int)x[k] > (int)a ? x[k] + k : b;             // ISETP + SEL(+add)
                if (MODE == 7) x[k] = x[k] * 3u + a;                                 // IMAD
            }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= x[k];
    if (s == 0x7fffffffu) out[0] = (int)s;
}

#endif  // !SB_HOST_EMUL

}  // namespace sb
