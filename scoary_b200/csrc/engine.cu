// engine.cu -- host side of libscoary_b200.so: context, device memory, the tree
// compiler and the C-ABI declared in include/scoary_b200.h.
//
// Data layout in HBM:
//   genes     uint64 [G][W]          isolate-column order, 16-byte row pitch
//   trait t   uint64 [W] value, [W] mask
//   lut       double2 [N+1]          log k! as (hi, lo)
//   per tree  uint32 [W32p][Gs]      gene bits in walk order, transposed so the
//                                    walk kernel's per-thread loads coalesce
//             uint16 [n_ops]         the compiled stack program
//             uint32 [P][W32p]       permuted label vectors in walk order
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/scoary_b200.h"
#include "fisher.cuh"
#include "walk.cuh"
#include "tree_build.cuh"
#include "epilogue.cuh"

extern "C" void sb_build_logfact_dd(int32_t n, double *hi_lo);

namespace {

std::string g_create_error;

enum Cat { CAT_PACK = 0, CAT_FISHER, CAT_SHUFFLE, CAT_WALK, CAT_PERMUTE, CAT_REDUCE, CAT_TREE, CAT_EPILOGUE, CAT_N };

struct TimedEvent {
    cudaEvent_t start, stop;
    int cat;
};

struct TraitSlot {
    bool has_trait = false;
    std::vector<uint64_t> h_value, h_mask;
    uint64_t *d_value = nullptr, *d_mask = nullptr;   // into the context's trait arena (not owned)
    // tree
    bool has_tree = false;
    bool finalized = false;        // labels + genesT built for the current genes/trait/tree
    int32_t n_leaves = 0, n_internal = 0, W32 = 0, W32p = 0, depth = 0, shift = 0, n_ops = 0;
    int32_t lab_base = 0;                // first label word behind the program in the constant pool (walk.cuh)
    std::vector<int32_t> h_leaf_to_col, h_leaf_of_pos;
    std::vector<uint16_t> h_ops;
    int32_t *d_walk_col = nullptr, *d_leaf_of_pos = nullptr;
    uint32_t *d_labels_leaf = nullptr;   // [W32]  by leaf id
    uint32_t *d_labels0 = nullptr;       // [W32p] walk order, unpermuted
    uint32_t *d_genesT = nullptr;        // [W32p][Gs]
    size_t genesT_cap = 0;               // bytes allocated behind d_genesT (reused when large enough)
    bool genesT_valid = false;
    int64_t Gs = 0;
};

}  // namespace

struct sb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    std::string err;
    int sm_count = 0;
    int max_smem_optin = 0;
    // genes
    int64_t G = 0;
    int32_t N = 0, W = 0;
    uint64_t *d_genes = nullptr;
    bool own_genes = false;
    uint64_t *d_genes_buf = nullptr;   // the context's own allocation (reused across sb_set_genes calls)
    size_t genes_cap = 0;
    uint64_t *d_trait_arena = nullptr;  // [SB_MAX_TRAITS][2][W]: value, mask of every trait slot, so that a Fisher launch
                                        // finds the vectors of consecutive traits in one block
    // lut
    double2 *d_lut = nullptr;
    int32_t lut_n = -1;
    TraitSlot traits[SB_MAX_TRAITS];
    // scratch
    void *d_scratch[16] = {nullptr};
    size_t scratch_bytes[16] = {0};
    int32_t *h_pinned_counter = nullptr;
    unsigned long long *d_walks = nullptr;   // walks of rounds whose slot count lives on the device
    int permute_mode = 0;                    // 0 auto, 1 threads = genes, 2 threads = labellings (sb_set_permute_mode)
    // stats / profiling
    sb_stats_t stats;
    bool profiling = false;
    std::vector<TimedEvent> pending;
    std::vector<TimedEvent> free_events;
    int *d_peak_out = nullptr;
};

namespace {

#define SB_CUDA(ctx, call)                                                                      \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            char buf_[512];                                                                     \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                     __FILE__, __LINE__);                                                       \
            (ctx)->err = buf_;                                                                  \
            return SB_ERR_CUDA;                                                                 \
        }                                                                                       \
    } while (0)

int fail(sb_ctx *ctx, int code, const char *msg)
{
    ctx->err = msg;
    return code;
}

int ensure_scratch(sb_ctx *ctx, int slot, size_t bytes)
{
    if (bytes <= ctx->scratch_bytes[slot]) return SB_OK;
    if (ctx->d_scratch[slot]) SB_CUDA(ctx, cudaFree(ctx->d_scratch[slot]));
    ctx->d_scratch[slot] = nullptr;
    ctx->scratch_bytes[slot] = 0;
    size_t cap = bytes + bytes / 8 + 256;
    SB_CUDA(ctx, cudaMalloc(&ctx->d_scratch[slot], cap));
    ctx->scratch_bytes[slot] = cap;
    return SB_OK;
}

// ---- profiling: a pair of events per launch, resolved lazily in sb_stats
struct Timed {
    sb_ctx *ctx;
    TimedEvent ev;
    bool on;
    Timed(sb_ctx *c, int cat) : ctx(c), on(c->profiling)
    {
        if (!on) return;
        if (!ctx->free_events.empty()) {
            ev = ctx->free_events.back();
            ctx->free_events.pop_back();
        } else {
            cudaEventCreate(&ev.start);
            cudaEventCreate(&ev.stop);
        }
        ev.cat = cat;
        cudaEventRecord(ev.start, ctx->stream);
    }
    ~Timed()
    {
        if (!on) return;
        cudaEventRecord(ev.stop, ctx->stream);
        ctx->pending.push_back(ev);
    }
};

void resolve_events(sb_ctx *ctx)
{
    for (auto &e : ctx->pending) {
        cudaEventSynchronize(e.stop);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.start, e.stop);
        switch (e.cat) {
            case CAT_PACK: ctx->stats.ms_pack += ms; break;
            case CAT_FISHER: ctx->stats.ms_fisher += ms; break;
            case CAT_SHUFFLE: ctx->stats.ms_shuffle += ms; break;
            case CAT_WALK: ctx->stats.ms_walk += ms; break;
            case CAT_PERMUTE: ctx->stats.ms_permute += ms; ctx->stats.launches_permute += 1; break;
            case CAT_REDUCE: ctx->stats.ms_reduce += ms; break;
            case CAT_TREE: ctx->stats.ms_pack += ms; break;   // tree construction is accounted with packing
            case CAT_EPILOGUE: ctx->stats.ms_epilogue += ms; break;
        }
        ctx->free_events.push_back(e);
    }
    ctx->pending.clear();
}

void free_tree(TraitSlot &s)
{
    s.h_ops.clear();
    cudaFree(s.d_walk_col); s.d_walk_col = nullptr;
    cudaFree(s.d_leaf_of_pos); s.d_leaf_of_pos = nullptr;
    cudaFree(s.d_labels_leaf); s.d_labels_leaf = nullptr;
    cudaFree(s.d_labels0); s.d_labels0 = nullptr;
    s.genesT_valid = false;              // the allocation itself is kept for the next tree
    s.has_tree = false;
    s.finalized = false;
}

void free_tree_storage(TraitSlot &s)
{
    free_tree(s);
    cudaFree(s.d_genesT); s.d_genesT = nullptr; s.genesT_cap = 0;
}

void free_trait(TraitSlot &s)
{
    s.d_value = nullptr;
    s.d_mask = nullptr;
    s.has_trait = false;
    s.finalized = false;
}

// ---- tree compiler --------------------------------------------------------
// Turns the flattened binary tree into the stack program described in walk.cuh.
struct Program {
    std::vector<uint16_t> ops;
    std::vector<int32_t> leaf_of_pos;   // walk position -> leaf id
    int depth = 0;                      // packed 16-bit entries pending at most (shared-memory stack units)
    int depth32 = 0;                    // 32-bit entries pending at most (per-thread local memory, <= WALK_STACK32)
};

// SB_WALK_PADDED: give every leaf-consuming op a place inside ONE leaf window.  Leaf runs are split at the window
// boundaries; a cherry op (2 + count leaves, no continuation) that does not fit the rest of its window starts the
// next one, and the positions it skips become padding (leaf_of_pos = -1: gene bit 0, no label).  The kernels apply
// the same rule ("does not fit -> open the next window"), so no flag is stored in the program.
bool pad_stream(Program &prog, std::string &err)
{
    const int W = sb::WALK_WINDOW;
    std::vector<uint16_t> ops;
    std::vector<int32_t> pos_leaf;
    ops.reserve(prog.ops.size() + prog.ops.size() / 4);
    pos_leaf.reserve(prog.leaf_of_pos.size() + prog.leaf_of_pos.size() / 4);
    size_t next = 0;          // next entry of the unpadded leaf order
    int room = 0;             // leaves left in the current window, as the kernel counts them
    auto take = [&](int n) {
        for (int k = 0; k < n; ++k) pos_leaf.push_back(prog.leaf_of_pos[next++]);
        room -= n;
    };
    for (const uint16_t op : prog.ops) {
        const int type = op & sb::OP_TYPE_MASK, cnt = op >> sb::OP_TYPE_BITS;
        const bool cherry = type == sb::OP_CHERRY_A16 || type == sb::OP_PUSH_CHERRY_A16 || type == sb::OP_CHERRY_B16 ||
                            type == sb::OP_CHERRY_B16_MERGE || type == (sb::OP_CHERRY_B16_MERGE | sb::OP_FLAG_BARE);
        if (cherry) {
            const int n = cnt + 2;
            if (n > W) { err = "internal error: cherry op longer than a leaf window"; return false; }
            if (n > room) {
                for (int k = 0; k < room; ++k) pos_leaf.push_back(-1);
                room = W;
            }
            take(n);
            ops.push_back(op);
        } else if (type == sb::OP_LEAF_A16 || type == sb::OP_LEAF_A32) {
            int left_ = cnt;
            while (left_ > 0) {
                if (room == 0) room = W;
                const int n = std::min(left_, room);
                take(n);
                ops.push_back((uint16_t)((n << sb::OP_TYPE_BITS) | type));
                left_ -= n;
            }
        } else {
            ops.push_back(op);
        }
    }
    if (next != prog.leaf_of_pos.size()) { err = "internal error: leaf stream and program disagree"; return false; }
    prog.ops.swap(ops);
    prog.leaf_of_pos.swap(pos_leaf);
    return true;
}

bool compile_tree(const int32_t *left, const int32_t *right, int32_t n_internal, Program &out, std::string &err)
{
    const int32_t n_leaves = n_internal + 1;
    // size[v]  leaves below v;  small[v]  size <= WALK_LIM16 (v is evaluated with packed 16-bit keys)
    // cat[v]   v is a "caterpillar" (a cherry plus single leaves): needs one accumulator only
    // need[v]  shared-memory stack units (10 words per gene pair; a 32-bit entry takes 2) used
    //          while v is evaluated into accumulator A
    std::vector<int32_t> need(n_internal, 0), size(n_internal, 0);
    std::vector<uint8_t> cat(n_internal, 0), first_left(n_internal, 1), second_in_b(n_internal, 0);
    std::vector<uint8_t> seen_leaf(n_leaves, 0), seen_node(n_internal, 0);
    auto small = [&](int32_t v) { return size[v] <= sb::WALK_LIM16; };
    for (int32_t v = 0; v < n_internal; ++v) {
        const int32_t ch[2] = {left[v], right[v]};
        for (int k = 0; k < 2; ++k) {
            if (ch[k] >= 0) {
                if (ch[k] >= v) { err = "tree nodes must list children before parents"; return false; }
                if (seen_node[ch[k]]) { err = "internal node referenced twice"; return false; }
                seen_node[ch[k]] = 1;
            } else {
                const int32_t leaf = ~ch[k];
                if (leaf < 0 || leaf >= n_leaves) { err = "leaf id out of range"; return false; }
                if (seen_leaf[leaf]) { err = "leaf referenced twice"; return false; }
                seen_leaf[leaf] = 1;
            }
        }
        const bool li = left[v] >= 0, ri = right[v] >= 0;
        size[v] = (li ? size[left[v]] : 1) + (ri ? size[right[v]] : 1);
        if (li && ri) {
            const int32_t a = left[v], b = right[v];
            auto cost = [&](int32_t f, int32_t s2, bool &in_b) {
                // padded streams: a B caterpillar must fit one leaf window (it has no continuation op)
                in_b = cat[s2] && small(s2) && (!sb::WALK_PADDED || size[s2] <= sb::WALK_WINDOW);
                return in_b ? need[f] : std::max(need[f], (small(f) ? 1 : 2) + need[s2]);
            };
            bool b_ab, b_ba;
            const int cost_ab = cost(a, b, b_ab), cost_ba = cost(b, a, b_ba);
            // cheaper stack first; on a tie prefer the order whose second child fits accumulator B
            const bool a_first = (cost_ab < cost_ba) || (cost_ab == cost_ba && (b_ab || !b_ba));
            first_left[v] = a_first;
            second_in_b[v] = a_first ? b_ab : b_ba;
            need[v] = a_first ? cost_ab : cost_ba;
        } else if (li) { need[v] = need[left[v]]; cat[v] = cat[left[v]]; }
        else if (ri) { need[v] = need[right[v]]; cat[v] = cat[right[v]]; }
        else { need[v] = 0; cat[v] = 1; }
    }
    for (int32_t v = 0; v + 1 < n_internal; ++v)
        if (!seen_node[v]) { err = "tree is not connected (root must be the last node)"; return false; }
    for (int32_t k = 0; k < n_leaves; ++k)
        if (!seen_leaf[k]) { err = "leaf missing from tree"; return false; }

    // emit one raw op per step (run-length encoded afterwards)
    std::vector<uint8_t> raw;
    raw.reserve((size_t)n_internal * 2);
    out.leaf_of_pos.clear();
    out.leaf_of_pos.reserve(n_leaves);
    struct Frame { int32_t node; int32_t phase; uint8_t into_b; };
    std::vector<Frame> st;
    st.push_back({n_internal - 1, 0, 0});
    while (!st.empty()) {
        Frame &f = st.back();
        const int32_t v = f.node;
        const bool li = left[v] >= 0, ri = right[v] >= 0;
        const uint8_t into_b = f.into_b;
        if (!li && !ri) {
            out.leaf_of_pos.push_back(~left[v]);
            out.leaf_of_pos.push_back(~right[v]);
            raw.push_back(into_b ? sb::OP_CHERRY_B16 : sb::OP_CHERRY_A16);
            st.pop_back();
        } else if (li != ri) {
            const int32_t inner = li ? left[v] : right[v];
            const int32_t leaf = li ? ~right[v] : ~left[v];
            if (f.phase == 0) {
                f.phase = 1;
                st.push_back({inner, 0, into_b});
            } else {
                out.leaf_of_pos.push_back(leaf);
                if (small(v)) {
                    raw.push_back(into_b ? sb::RAW_LEAF_B16 : sb::OP_LEAF_A16);
                } else {
                    if (small(inner)) raw.push_back(sb::OP_WIDEN_A);
                    raw.push_back(sb::OP_LEAF_A32);
                }
                st.pop_back();
            }
        } else {
            const int32_t first = first_left[v] ? left[v] : right[v];
            const int32_t second = first_left[v] ? right[v] : left[v];
            const bool in_b = second_in_b[v] != 0;
            if (f.phase == 0) {
                f.phase = 1;
                st.push_back({first, 0, 0});
            } else if (f.phase == 1) {
                f.phase = 2;
                if (!in_b) raw.push_back(small(first) ? sb::OP_PUSH16 : sb::OP_PUSH32);
                st.push_back({second, 0, (uint8_t)(in_b ? 1 : 0)});
            } else {
                if (in_b) {
                    if (small(v)) raw.push_back(sb::RAW_MERGE_AB16);
                    else {
                        if (small(first)) raw.push_back(sb::OP_WIDEN_A);
                        raw.push_back(sb::OP_MERGE_A32_B16);
                    }
                } else if (small(v)) {
                    raw.push_back(sb::OP_MERGE_POP16);
                } else {
                    if (small(second)) raw.push_back(sb::OP_WIDEN_A);
                    raw.push_back(small(first) ? sb::OP_MERGE_POPW : sb::OP_MERGE_POP32);
                }
                st.pop_back();
            }
        }
    }
    if (small(n_internal - 1)) raw.push_back(sb::OP_WIDEN_A);   // the program always ends in 32-bit mode
    // peephole fusion + run-length encoding:
    //   [PUSH16] CHERRY_A16 LEAF_A16*            -> (PUSH_)CHERRY_A16(n)
    //   CHERRY_B16 LEAF_B16* [MERGE_AB16]        -> CHERRY_B16(_MERGE)(n)
    //   LEAF_A16+ / LEAF_A32+ / MERGE_POP*+      -> one op with a count
    out.ops.clear();
    int depth = 0, sp = 0, depth32 = 0, sp32 = 0;   // the two stacks of walk.cuh: 16-bit entries, 32-bit entries
    auto emit = [&](int kind, size_t cnt) {
        out.ops.push_back((uint16_t)((cnt << sb::OP_TYPE_BITS) | (unsigned)kind));
    };
    size_t i = 0;
    const size_t n_raw = raw.size();
    while (i < n_raw) {
        const uint8_t kind = raw[i];
        if (kind == sb::OP_PUSH16 || kind == sb::OP_CHERRY_A16) {
            const bool push = kind == sb::OP_PUSH16;
            if (push) { sp += 1; depth = std::max(depth, sp); }
            if (push && !(i + 1 < n_raw && raw[i + 1] == sb::OP_CHERRY_A16)) {   // push not followed by a cherry
                emit(sb::OP_PUSH16, 1);
                ++i;
                continue;
            }
            size_t j = i + (push ? 2 : 1);
            size_t leaves = 0;
            const size_t cap = sb::WALK_PADDED ? (size_t)(sb::WALK_WINDOW - 2) : (size_t)sb::OP_MAX_COUNT;
            while (j < n_raw && raw[j] == sb::OP_LEAF_A16 && leaves < cap) { ++j; ++leaves; }
            emit(push ? sb::OP_PUSH_CHERRY_A16 : sb::OP_CHERRY_A16, leaves);
            i = j;
        } else if (kind == sb::OP_CHERRY_B16) {
            size_t j = i + 1, leaves = 0;
            while (j < n_raw && raw[j] == sb::RAW_LEAF_B16 && leaves < (size_t)sb::OP_MAX_COUNT) { ++j; ++leaves; }
            if (j < n_raw && raw[j] == sb::RAW_LEAF_B16) { err = "caterpillar too long for one op"; return false; }
            const bool merge = j < n_raw && raw[j] == sb::RAW_MERGE_AB16;
            emit(merge ? (sb::OP_CHERRY_B16_MERGE | (leaves == 0 ? sb::OP_FLAG_BARE : 0)) : sb::OP_CHERRY_B16, leaves);
            i = j + (merge ? 1 : 0);
        } else if (kind == sb::RAW_LEAF_B16 || kind == sb::RAW_MERGE_AB16) {
            err = "internal error: B-accumulator step outside a B evaluation";
            return false;
        } else {
            size_t j = i + 1;
            // (a packed pop is one op each: as a counted loop the kernel shuffled 16 registers around every merge)
            const bool runs = (kind == sb::OP_LEAF_A16 || kind == sb::OP_LEAF_A32 || kind == sb::OP_MERGE_POP32 ||
                               kind == sb::OP_MERGE_POPW);
            if (runs) while (j < n_raw && raw[j] == kind) ++j;
            size_t cnt = j - i;
            if (kind == sb::OP_PUSH32) { sp32 += 1; depth32 = std::max(depth32, sp32); }
            if (kind == sb::OP_MERGE_POP16 || kind == sb::OP_MERGE_POPW) sp -= (int)cnt;
            if (kind == sb::OP_MERGE_POP32) sp32 -= (int)cnt;
            while (cnt > 0) {
                const size_t c = std::min<size_t>(cnt, (size_t)sb::OP_MAX_COUNT);
                emit(kind, runs ? c : 1);
                cnt -= c;
            }
            i = j;
        }
    }
    out.ops.push_back((uint16_t)sb::OP_END);
    if (sp != 0 || sp32 != 0) { err = "internal error: unbalanced stack program"; return false; }
    if (depth32 > sb::WALK_STACK32) { err = "tree too deep for the 32-bit DP stack"; return false; }
    out.depth = depth;   // stack units of 10 words per gene pair
    out.depth32 = depth32;
    if (sb::WALK_PADDED && !pad_stream(out, err)) return false;
    return true;
}

int ensure_lut(sb_ctx *ctx, int32_t n)
{
    if (n <= ctx->lut_n) return SB_OK;
    std::vector<double> h((size_t)(n + 1) * 2);
    sb_build_logfact_dd(n, h.data());
    if (ctx->d_lut) SB_CUDA(ctx, cudaFree(ctx->d_lut));
    ctx->d_lut = nullptr;
    SB_CUDA(ctx, cudaMalloc(&ctx->d_lut, sizeof(double2) * (size_t)(n + 1)));
    SB_CUDA(ctx, cudaMemcpyAsync(ctx->d_lut, h.data(), sizeof(double2) * (size_t)(n + 1), cudaMemcpyHostToDevice,
                                 ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // h goes out of scope
    ctx->stats.h2d_bytes += (int64_t)sizeof(double2) * (n + 1);
    ctx->lut_n = n;
    return SB_OK;
}

int set_genes_common(sb_ctx *ctx, int64_t G, int32_t N, int32_t W)
{
    if (G <= 0 || N <= 0) return fail(ctx, SB_ERR_ARG, "sb_set_genes: G and N must be positive");
    if (G > (int64_t)0x7fffff00) return fail(ctx, SB_ERR_ARG, "sb_set_genes: at most 2^31 - 256 gene rows per context");
    if (W < (N + 63) / 64 || (W & 1)) return fail(ctx, SB_ERR_ARG, "sb_set_genes: W must be even and >= ceil(N/64)");
    ctx->d_genes = nullptr;
    ctx->own_genes = false;
    if (N != ctx->N || W != ctx->W || !ctx->d_trait_arena) {
        for (auto &s : ctx->traits) { free_trait(s); free_tree(s); }
        if (ctx->d_trait_arena) SB_CUDA(ctx, cudaFree(ctx->d_trait_arena));
        ctx->d_trait_arena = nullptr;
        SB_CUDA(ctx, cudaMalloc(&ctx->d_trait_arena, sizeof(uint64_t) * SB_MAX_TRAITS * 2 * (size_t)W));
    }
    for (auto &s : ctx->traits) {   // gene-dependent derived data is stale
        s.finalized = false;
        s.genesT_valid = false;
    }
    ctx->G = G; ctx->N = N; ctx->W = W;
    return ensure_lut(ctx, N);
}

// Build labels (by leaf id and in walk order) and the transposed gene matrix for slot t.
int finalize_slot(sb_ctx *ctx, int32_t t)
{
    TraitSlot &s = ctx->traits[t];
    if (!ctx->d_genes) return fail(ctx, SB_ERR_STATE, "genes not set (sb_set_genes)");
    if (!s.has_trait) return fail(ctx, SB_ERR_STATE, "trait not set (sb_set_trait)");
    if (!s.has_tree) return fail(ctx, SB_ERR_STATE, "tree not set (sb_set_tree)");
    if (s.finalized) return SB_OK;
    std::vector<uint32_t> lab_leaf(s.W32, 0u), lab0(s.W32p, 0u);
    for (int32_t k = 0; k < s.n_leaves; ++k) {
        const int32_t col = s.h_leaf_to_col[k];
        if (col < 0 || col >= ctx->N) return fail(ctx, SB_ERR_ARG, "sb_set_tree: leaf_to_col out of range");
        if (!((s.h_mask[col >> 6] >> (col & 63)) & 1ULL))
            return fail(ctx, SB_ERR_ARG, "tree has a leaf whose trait value is missing: prune the tree first "
                                         "(PruneForMissing, scoary/methods.py:709-739)");
        if ((s.h_value[col >> 6] >> (col & 63)) & 1ULL) lab_leaf[k >> 5] |= (1u << (k & 31));
    }
    for (int32_t pos = 0; pos < (int32_t)s.h_leaf_of_pos.size(); ++pos) {
        const int32_t leaf = s.h_leaf_of_pos[pos];
        if (leaf < 0) continue;      // pad position (SB_WALK_PADDED)
        if ((lab_leaf[leaf >> 5] >> (leaf & 31)) & 1u) lab0[pos >> 5] |= (1u << (pos & 31));
    }
    if (!s.d_labels_leaf) SB_CUDA(ctx, cudaMalloc(&s.d_labels_leaf, sizeof(uint32_t) * s.W32));
    if (!s.d_labels0) SB_CUDA(ctx, cudaMalloc(&s.d_labels0, sizeof(uint32_t) * s.W32p));
    SB_CUDA(ctx, cudaMemcpyAsync(s.d_labels_leaf, lab_leaf.data(), sizeof(uint32_t) * s.W32, cudaMemcpyHostToDevice,
                                 ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(s.d_labels0, lab0.data(), sizeof(uint32_t) * s.W32p, cudaMemcpyHostToDevice,
                                 ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.h2d_bytes += (int64_t)sizeof(uint32_t) * (s.W32 + s.W32p);
    if (!s.genesT_valid) {
        s.Gs = (ctx->G + 31) / 32 * 32;
        const size_t need = sizeof(uint32_t) * (size_t)s.W32p * (size_t)s.Gs;
        if (need > s.genesT_cap) {
            if (s.d_genesT) SB_CUDA(ctx, cudaFree(s.d_genesT));
            s.d_genesT = nullptr; s.genesT_cap = 0;
            SB_CUDA(ctx, cudaMalloc(&s.d_genesT, need));
            s.genesT_cap = need;
        }
        const size_t smem = sizeof(uint32_t) * 32 * (size_t)(2 * ctx->W + 1);
        if (smem > (size_t)ctx->max_smem_optin) return fail(ctx, SB_ERR_ARG, "too many isolates for the pack kernel");
        SB_CUDA(ctx, cudaFuncSetAttribute(sb::pack_walk_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
        Timed tm(ctx, CAT_PACK);
        const unsigned blocks = (unsigned)((ctx->G + 31) / 32);
        sb::pack_walk_order_kernel<<<blocks, 256, smem, ctx->stream>>>(ctx->d_genes, ctx->G, ctx->W, s.d_walk_col,
                                                                       (int)s.h_leaf_of_pos.size(), s.W32p, s.Gs,
                                                                       s.d_genesT);
        ctx->stats.kernel_launches += 1;
        SB_CUDA(ctx, cudaGetLastError());
        s.genesT_valid = true;
    }
    s.finalized = true;
    return SB_OK;
}

// Contingency tables + Fisher p for traits t0 .. t0 + nT - 1, every gene row read once per launch of up to
// FISHER_MAX_TRAITS traits.  Outputs are trait-major: counts [nT][G][4], p [nT][G], hash [nT][G][2].
int launch_fisher(sb_ctx *ctx, int32_t t0, int32_t nT, int32_t *d_counts, double *d_p, uint64_t *d_hash)
{
    if (nT < 1 || t0 < 0 || t0 + nT > SB_MAX_TRAITS) return fail(ctx, SB_ERR_ARG, "trait index out of range");
    if (!ctx->d_genes) return fail(ctx, SB_ERR_STATE, "genes not set (sb_set_genes)");
    for (int32_t t = t0; t < t0 + nT; ++t)
        if (!ctx->traits[t].has_trait) return fail(ctx, SB_ERR_STATE, "trait not set (sb_set_trait)");
    const size_t row_bytes = (size_t)ctx->W * 8;
    const size_t lut_bytes = sizeof(double2) * (size_t)(ctx->lut_n + 1);
    constexpr size_t ROWS_PER_WARP = sb::F_GENES;
    const bool hash = d_hash != nullptr;
    const int64_t G = ctx->G;
    for (int32_t c0 = 0; c0 < nT; c0 += sb::FISHER_MAX_TRAITS) {
        const int32_t n = std::min<int32_t>(sb::FISHER_MAX_TRAITS, nT - c0);
        sb::FisherArgs A;
        A.genes = ctx->d_genes; A.G = G; A.W = ctx->W; A.Wn = (ctx->N + 63) / 64;
        A.traits = ctx->d_trait_arena + (size_t)(t0 + c0) * 2 * ctx->W; A.n_traits = n;
        A.lut = ctx->d_lut; A.lut_n = ctx->lut_n;
        A.counts = d_counts ? d_counts + (size_t)c0 * G * 4 : nullptr;
        A.p = d_p ? d_p + (size_t)c0 * G : nullptr;
        A.hash = d_hash ? d_hash + (size_t)c0 * G * 2 : nullptr;
        // [2 mbarriers per warp][per trait: value & mask, mask][2 buffers of 4 rows per warp][LUT if it fits];
        // long rows leave room for fewer warps (N = 32 766: 4 KB rows, 6 warps)
        const size_t budget = (size_t)ctx->max_smem_optin - 256;     // minus the kernel's static shared memory
        const size_t per_warp = 16 + 2 * ROWS_PER_WARP * row_bytes, traits_bytes = 2 * row_bytes * n;
        if (traits_bytes + per_warp > budget) return fail(ctx, SB_ERR_ARG, "row too long for shared memory");
        const size_t NW = std::min<size_t>(sb::FISHER_THREADS / 32, (budget - traits_bytes) / per_warp);
        const size_t fixed = 16 * NW + traits_bytes + 2 * NW * ROWS_PER_WARP * row_bytes;
        const bool lut_smem = fixed + lut_bytes <= budget;
        const size_t smem = fixed + (lut_smem ? lut_bytes : 0);
        const int64_t rows_per_cta = (int64_t)(NW * ROWS_PER_WARP);
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((G + rows_per_cta - 1) / rows_per_cta, ctx->sm_count));
        Timed tm(ctx, CAT_FISHER);
#define SB_LAUNCH_FISHER(L, H)                                                                                     \
    do {                                                                                                           \
        SB_CUDA(ctx, cudaFuncSetAttribute(sb::fisher_kernel<L, H>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                          (int)smem));                                                             \
        sb::fisher_kernel<L, H><<<grid, (unsigned)(NW * 32), smem, ctx->stream>>>(A);                              \
    } while (0)
        if (lut_smem && hash) SB_LAUNCH_FISHER(true, true);
        else if (lut_smem) SB_LAUNCH_FISHER(true, false);
        else if (hash) SB_LAUNCH_FISHER(false, true);
        else SB_LAUNCH_FISHER(false, false);
#undef SB_LAUNCH_FISHER
        ctx->stats.kernel_launches += 1;
        ctx->stats.tests_contingency += G * n;
        SB_CUDA(ctx, cudaGetLastError());
    }
    return SB_OK;
}

// DP stack: s.depth pushes per thread; a push holds all packed states of the thread (10 words each with both
// passes, K4; 5 with one, K5), padded to whole 128-bit chunks (sb::walk_push_words)
size_t walk_smem_bytes(const TraitSlot &s, bool dual)
{
    return sizeof(int) * (size_t)sb::walk_push_words(dual, dual ? 1 : sb::WALK_NLAB) * (size_t)std::max(1, s.depth) *
           sb::WALK_THREADS;
}

void fill_walk_args(const TraitSlot &s, sb::WalkArgs &A, const int64_t *d_gene_idx, int64_t S)
{
    memset(&A, 0, sizeof A);
    A.genesT = s.d_genesT; A.Gs = s.Gs; A.gene_idx = d_gene_idx; A.S = S; A.S_total = S; A.slot_idx = nullptr;
    A.W32p = s.W32p; A.shift = s.shift; A.lab_base = s.lab_base; A.tile_threads = sb::WALK_THREADS;
}

// label vectors (or, in transposed launches, gene rows) that fit behind the program of slot s
int label_capacity(const TraitSlot &s) { return (sb::C_POOL_WORDS - s.lab_base) / s.W32p; }

// the compiled program of slot s -> the start of the constant pool (stream ordered; sb_set_tree checked the size)
int upload_program(sb_ctx *ctx, const TraitSlot &s)
{
    SB_CUDA(ctx, cudaMemcpyToSymbolAsync(sb::c_pool, s.h_ops.data(), sizeof(uint16_t) * s.h_ops.size(), 0,
                                         cudaMemcpyHostToDevice, ctx->stream));
    return SB_OK;
}

// n vectors of W32p words (device memory) -> the label area of the constant pool
int upload_labels(sb_ctx *ctx, const TraitSlot &s, const uint32_t *d_src, int n)
{
    SB_CUDA(ctx, cudaMemcpyToSymbolAsync(sb::c_pool, d_src, sizeof(uint32_t) * (size_t)n * s.W32p,
                                         sizeof(uint32_t) * (size_t)s.lab_base, cudaMemcpyDeviceToDevice, ctx->stream));
    return SB_OK;
}

int launch_pairwise(sb_ctx *ctx, int32_t t, const int64_t *d_gene_idx, int64_t S, int32_t *d_pairs)
{
    TraitSlot &s = ctx->traits[t];
    int rc = upload_program(ctx, s);
    if (rc) return rc;
    rc = upload_labels(ctx, s, s.d_labels0, 1);
    if (rc) return rc;
    sb::WalkArgs A;
    fill_walk_args(s, A, d_gene_idx, S);
    A.pairs = d_pairs;
    const size_t smem = walk_smem_bytes(s, true);
    if (smem > (size_t)ctx->max_smem_optin) return fail(ctx, SB_ERR_ARG, "tree too deep for shared memory");
    SB_CUDA(ctx, cudaFuncSetAttribute(sb::walk_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Timed tm(ctx, CAT_WALK);
    const int64_t per_block = (int64_t)sb::WALK_THREADS * sb::WALK_NP;
    dim3 grid((unsigned)((S + per_block - 1) / per_block), 1, 1);
    sb::walk_pairs_kernel<<<grid, sb::WALK_THREADS, smem, ctx->stream>>>(A);
    ctx->stats.kernel_launches += 1;
    ctx->stats.tests_walks += S;
    SB_CUDA(ctx, cudaGetLastError());
    return SB_OK;
}

int launch_shuffle(sb_ctx *ctx, int32_t t, int32_t P, int32_t perm_first, uint64_t seed, uint32_t *d_labelsW, uint8_t *d_dbg)
{
    TraitSlot &s = ctx->traits[t];
    int T = 64;       // permutations (threads) per block: 64 label vectors in shared memory, 32 above ~28 000 leaves
    while (T > 8 && sizeof(uint32_t) * (size_t)s.W32 * T > (size_t)ctx->max_smem_optin) T /= 2;
    const size_t smem = sizeof(uint32_t) * (size_t)s.W32 * T;
    if (smem > (size_t)ctx->max_smem_optin) return fail(ctx, SB_ERR_ARG, "too many leaves for the shuffle kernel");
    SB_CUDA(ctx, cudaFuncSetAttribute(sb::shuffle_labels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Timed tm(ctx, CAT_SHUFFLE);
    sb::shuffle_labels_kernel<<<(P + T - 1) / T, T, smem, ctx->stream>>>(s.d_labels_leaf, s.n_leaves, s.W32, s.W32p,
                                                                       s.d_leaf_of_pos, seed, t, P, perm_first, d_labelsW, d_dbg);
    ctx->stats.kernel_launches += 1;
    SB_CUDA(ctx, cudaGetLastError());
    return SB_OK;
}

// ---- K5 launch plans ---------------------------------------------------------
// How well a launch shape fills the GPU: lanes that carry a walk x resident block slots that get a block.
double fill_fraction(int64_t threads_domain, int64_t blocks_y, int slots)
{
    const int64_t per_block = (int64_t)sb::WALK_THREADS * sb::WALK_NP;
    const int64_t tiles = (threads_domain + per_block - 1) / per_block;
    const double lanes = (double)threads_domain / (double)(tiles * per_block);
    const double blocks = (double)(tiles * blocks_y);
    return lanes * std::min(1.0, blocks / (double)slots);
}

// Labellings per launch (n <= cap) and per block (ppi) for `tiles` thread tiles on `slots` resident blocks.  Every block
// of a launch does the same work, so a launch runs in waves of `slots` blocks.  Measured (tools/rule_probe.py,
// profiles/r2_rule_probe.txt; 592 slots): a full wave takes 1.13 ms, a wave filled to a fraction f still takes
// (0.25 + 0.75 f) of that -- blocks that share an SM with fewer others run faster, but not proportionally -- so whole
// waves are cheapest per labelling: cost = (floor(w) + (0.25 + 0.75 frac(w) if frac(w) > 0) ) x ppi + 0.05, w = blocks / slots.
// 5 700 genes (8 tiles of 768) on 592 slots: 74 labellings x 1 per block = 592 blocks, exactly one wave, 155 ms per
// 10 000 permutations where the largest launch the constant pool allows (91: 1.23 waves) takes 168 ms.  (Figures of
// the 192 x 4 block shape the model was measured with; the rule itself does not depend on the shape.)
void plan_launch(int64_t tiles, int cap, int slots, int ppi_min, int ppi_max, int *ppi_out, int *n_out)
{
    double best = 1e300;
    int best_ppi = ppi_min, best_n = std::max(1, cap);
    for (int ppi = ppi_max; ppi >= ppi_min; ppi /= 2) {
        for (int n = cap / ppi * ppi; n >= ppi; n -= ppi) {
            const double w = (double)(tiles * (n / ppi)) / (double)slots;
            const double full = std::floor(w + 1e-9), frac = w - full;
            const double waves = full + (frac > 1e-9 ? 0.25 + 0.75 * frac : 0.0);
            const double cost = (waves * ppi + 0.05) / n * (1.0 + 0.004 * (ppi_max / ppi - 1));
            if (cost < best) { best = cost; best_ppi = ppi; best_n = n; }
        }
        if (ppi == 1) break;
    }
    *ppi_out = best_ppi;
    *n_out = best_n;
}

// Transposed launches (walk.cuh): threads = labellings, constant rows = the genes of result slots list[e] / e.
// Walks labellings [perm_lo, perm_hi) for n_rows genes; hit bytes go to d_hits [slot][Ps].
int launch_rows(sb_ctx *ctx, const TraitSlot &s, const uint32_t *d_labelsT, int64_t Ps, int perm_lo, int perm_hi,
                const int64_t *d_gene_idx, const int32_t *d_list, int64_t n_rows, const int32_t *d_unperm,
                uint8_t *d_hits, uint32_t *d_rowsW)
{
    if (n_rows <= 0 || perm_hi <= perm_lo) return SB_OK;
    {
        const int64_t n = n_rows * s.W32p;
        sb::gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(s.d_genesT, s.Gs, s.W32p, d_gene_idx,
                                                                                    d_list, (int)n_rows, d_rowsW);
        ctx->stats.kernel_launches += 1;
        SB_CUDA(ctx, cudaGetLastError());
    }
    const int cap = label_capacity(s);
    const int64_t Pn = perm_hi - perm_lo;
    const int64_t per_block = (int64_t)sb::WALK_THREADS * sb::WALK_NP;
    const int64_t tiles = (Pn + per_block - 1) / per_block;
    int ppi = sb::PERMS_PER_ITEM_MAX;      // rows per block: fewer when that is what it takes to give every SM work
    while (ppi > 1 && tiles * ((std::min<int64_t>(cap, n_rows) + ppi - 1) / ppi) < 2LL * 7 * ctx->sm_count) ppi /= 2;
    const size_t smem = sizeof(int) * (size_t)sb::walk_push_words(false, 1) * (size_t)std::max(1, s.depth) * sb::WALK_THREADS;
    SB_CUDA(ctx, cudaFuncSetAttribute(sb::walk_permute_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int64_t base = 0; base < n_rows; base += cap) {
        const int n = (int)std::min<int64_t>(cap, n_rows - base);
        int rc = upload_labels(ctx, s, d_rowsW + (size_t)base * s.W32p, n);
        if (rc) return rc;
        sb::WalkArgs A;
        memset(&A, 0, sizeof A);
        A.genesT = d_labelsT + perm_lo; A.Gs = Ps; A.S = Pn; A.S_total = Ps;
        A.slot_idx = d_list; A.row_base = (int32_t)base;
        A.W32p = s.W32p; A.shift = s.shift; A.lab_base = s.lab_base; A.tile_threads = sb::WALK_THREADS;
        A.n_perms = n; A.ppi = ppi; A.items_per_tile = (n + ppi - 1) / ppi;
        A.unperm = d_unperm; A.hits = d_hits + perm_lo;
        dim3 grid((unsigned)tiles, (unsigned)A.items_per_tile, 1);
        Timed tm(ctx, CAT_PERMUTE);
        sb::walk_permute_kernel<true><<<grid, sb::WALK_THREADS, smem, ctx->stream>>>(A);
        ctx->stats.kernel_launches += 1;
        ctx->stats.tests_walks += (int64_t)n * Pn;
        SB_CUDA(ctx, cudaGetLastError());
    }
    return SB_OK;
}

int launch_permute_transposed(sb_ctx *ctx, TraitSlot &s, const int64_t *d_gene_idx, int64_t S, int32_t P,
                              int32_t early_stop, const int32_t *d_rmin, const int32_t *d_unperm,
                              const uint32_t *d_labelsW, int32_t *d_r, int32_t *d_n_done)
{
    const int64_t Ps = ((int64_t)P + 31) / 32 * 32;
    int rc = ensure_scratch(ctx, 9, sizeof(uint32_t) * (size_t)s.W32p * (size_t)Ps);
    if (rc) return rc;
    rc = ensure_scratch(ctx, 10, sizeof(uint32_t) * (size_t)s.W32p * (size_t)S);
    if (rc) return rc;
    rc = ensure_scratch(ctx, 1, (size_t)S * (size_t)Ps);
    if (rc) return rc;
    rc = ensure_scratch(ctx, 8, sizeof(int32_t) * ((size_t)S + 4));
    if (rc) return rc;
    uint32_t *d_labelsT = (uint32_t *)ctx->d_scratch[9], *d_rowsW = (uint32_t *)ctx->d_scratch[10];
    uint8_t *d_hits = (uint8_t *)ctx->d_scratch[1];
    int32_t *d_list = (int32_t *)ctx->d_scratch[8], *d_counter = d_list + S;
    {
        Timed tm(ctx, CAT_SHUFFLE);
        dim3 grid((unsigned)((P + 31) / 32), (unsigned)((s.W32p + 31) / 32), 1);
        sb::transpose_labels_kernel<<<grid, 256, 0, ctx->stream>>>(d_labelsW, P, s.W32p, Ps, d_labelsT);
        ctx->stats.kernel_launches += 1;
        SB_CUDA(ctx, cudaGetLastError());
    }
    auto reduce = [&](const int32_t *list, int64_t n, int n_avail) -> int {
        Timed tm(ctx, CAT_REDUCE);
        sb::reduce_hit_rows_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, ctx->stream>>>(
            d_hits, Ps, list, (int)n, n_avail, P, early_stop, d_rmin, d_r, d_n_done, d_list, d_counter);
        ctx->stats.kernel_launches += 1;
        SB_CUDA(ctx, cudaGetLastError());
        return SB_OK;
    };
    // Reference-rule mode: most genes stop within the first few dozen labellings (methods.py:1360-1363), so the
    // first tile of labellings is walked for everything and only the genes still running see the rest.
    const int first = early_stop ? (int)std::min<int64_t>(P, (int64_t)sb::WALK_THREADS * sb::WALK_NP) : P;
    rc = launch_rows(ctx, s, d_labelsT, Ps, 0, first, d_gene_idx, nullptr, S, d_unperm, d_hits, d_rowsW);
    if (rc) return rc;
    SB_CUDA(ctx, cudaMemsetAsync(d_counter, 0, sizeof(int32_t), ctx->stream));
    rc = reduce(nullptr, S, first);
    if (rc || first >= P) return rc;
    if (!ctx->h_pinned_counter) SB_CUDA(ctx, cudaMallocHost(&ctx->h_pinned_counter, sizeof(int32_t)));
    SB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned_counter, d_counter, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int64_t n_alive = *ctx->h_pinned_counter;
    if (n_alive == 0) return SB_OK;
    rc = launch_rows(ctx, s, d_labelsT, Ps, first, P, d_gene_idx, d_list, n_alive, d_unperm, d_hits, d_rowsW);
    if (rc) return rc;
    return reduce(d_list, n_alive, P);   // appends nothing: n_avail == P
}

// d_unperm: [S][3] device (input, already computed); d_r / d_n_done outputs
int launch_permute(sb_ctx *ctx, int32_t t, const int64_t *d_gene_idx, int64_t S, int32_t P, uint64_t seed,
                   int32_t early_stop, const int32_t *d_rmin, const int32_t *d_unperm, int32_t *d_r, int32_t *d_n_done,
                   int32_t perm_first = 0)
{
    TraitSlot &s = ctx->traits[t];
    // scratch 0: labelsW [P][W32p]; scratch 1: hit bytes of one slice; scratch 8: work lists + counters
    int rc = ensure_scratch(ctx, 0, sizeof(uint32_t) * (size_t)P * s.W32p);
    if (rc) return rc;
    uint32_t *d_labelsW = (uint32_t *)ctx->d_scratch[0];
    rc = launch_shuffle(ctx, t, P, perm_first, seed, d_labelsW, nullptr);
    if (rc) return rc;
    rc = upload_program(ctx, s);
    if (rc) return rc;

    const size_t smem = walk_smem_bytes(s, false);
    if (smem > (size_t)ctx->max_smem_optin) return fail(ctx, SB_ERR_ARG, "tree too deep for shared memory");
    SB_CUDA(ctx, cudaFuncSetAttribute(sb::walk_permute_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // Threads per block: WALK_THREADS, or fewer when that leaves the last tile of a SMALL work list nearly empty
    // (6 250 genes -- the north_star job on 8 GPUs -- are 12.2 tiles of 128 x 4 but 16.3 tiles of 96 x 4: 96 % of the
    // lanes carry a walk instead of 94 %).  Only exhaustive mode: reference-rule rounds shrink their lists anyway.
    int tile_threads = sb::WALK_THREADS;
    if (!early_stop && S < 24LL * sb::WALK_THREADS * sb::WALK_NP) {
        double best = 0.0;
        for (int tt = sb::WALK_THREADS; tt >= 64; tt -= 32) {
            const int64_t per = (int64_t)tt * sb::WALK_NP, tl = (S + per - 1) / per;
            const double util = (double)S / (double)(tl * per);
            if (util > best + 0.02) { best = util; tile_threads = tt; }
        }
    }
    int per_sm = 0;
    SB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sb::walk_permute_kernel<false>, tile_threads, smem));
    const int slots = std::max(1, per_sm) * ctx->sm_count;      // resident K5 blocks
    const int label_cap = label_capacity(s);
    // threads = genes (the default) or threads = labellings: whichever shape fills the GPU better
    bool transposed = false;
    if (sb::WALK_NLAB == 1 && ctx->permute_mode != 1) {
        if (!early_stop) {
            const double f_genes = fill_fraction(S, std::min(label_cap, P), slots);
            const double f_perms = fill_fraction(P, std::min<int64_t>(label_cap, S), slots);
            transposed = f_perms > 1.5 * f_genes;
        } else {
            // Reference-rule mode, estimated times (profiles/r2_few_genes_probe.jsonl, r2_rule_probe.txt): threads = genes
            // pays 64 labellings for every gene and then ~(P - 64) / 74 sequential rounds of at least one wave of blocks
            // each; threads = labellings pays the first tile of labellings for every gene and the rest for the genes
            // still running (unknown in advance: half of them assumed).
            const double rate = 4.0e5 * 5000.0 / std::max(64, s.n_leaves);        // walks per ms at this tree size
            const double per_tile = (double)sb::WALK_THREADS * sb::WALK_NP, frac = 0.5;
            const int n_round = std::max(1, std::min(label_cap, 74));
            const double t_wave = slots * per_tile / rate;
            const double waves = std::max(1.0, std::ceil(frac * S / per_tile) * n_round / slots);
            const double t_genes = S * 64.0 / rate + std::ceil(std::max(0, P - 64) / (double)n_round) * t_wave * waves;
            const double first = std::min<double>(P, per_tile);
            const double t_perms = S * first / rate + frac * S * (P - first) / rate + 0.5;
            transposed = t_perms < t_genes;
        }
        transposed = transposed || ctx->permute_mode == 2;
    }
    ctx->stats.calls_transposed += transposed ? 1 : 0;
    if (transposed)
        return launch_permute_transposed(ctx, s, d_gene_idx, S, P, early_stop, d_rmin, d_unperm, d_labelsW, d_r, d_n_done);

    // labellings per launch and per block, in whole waves of resident blocks (plan_launch).  Reference-rule mode
    // plans every round for the genes still running, one labelling per block (its hit rows hold one flag each).
    const int64_t per_block = (int64_t)tile_threads * sb::WALK_NP;
    const int64_t tiles_all = (S + per_block - 1) / per_block;
    int ppi = sb::WALK_NLAB, n_launch = std::min(label_cap, P);
    // one labelling per block: measured 3 % faster than two or four per block at equal wave counts (more, shorter
    // blocks balance better; the hit buffer is per launch, so its size no longer argues for more)
    if (!early_stop) plan_launch(tiles_all, std::min(label_cap, P), slots, sb::WALK_NLAB, sb::WALK_NLAB, &ppi, &n_launch);
    rc = ensure_scratch(ctx, 1, (size_t)((std::min(label_cap, P) + ppi - 1) / ppi) * (size_t)S);
    if (rc) return rc;
    uint8_t *d_hits = (uint8_t *)ctx->d_scratch[1];

    // one slice: labellings [base, base + n_perms) for the slots of d_list (n_bound of them at most; the exact
    // count is n_bound itself or, for rounds enqueued without a host round trip, *d_count), then the bookkeeping
    const uint32_t *d_genesC = nullptr;     // reference-rule mode: compacted columns of the slots still running
    const int32_t *d_col_of_slot = nullptr;
    int64_t genesC_pad = 0;
    auto slice = [&](int base, int n_perms, const int32_t *d_list, int64_t n_bound, const int32_t *d_count,
                     int32_t *d_list_out, int32_t *d_count_out) -> int {
        int rc2 = upload_labels(ctx, s, d_labelsW + (size_t)base * s.W32p, n_perms);
        if (rc2) return rc2;
        sb::WalkArgs A;
        fill_walk_args(s, A, d_gene_idx, n_bound);
        if (d_genesC) { A.genesT = d_genesC; A.Gs = genesC_pad; A.col_idx = d_col_of_slot; }
        A.S_total = S;
        A.S_dev = d_count;
        A.tile_threads = tile_threads;
        A.slot_idx = d_list;
        A.n_perms = n_perms;
        A.ppi = ppi;
        A.items_per_tile = (n_perms + ppi - 1) / ppi;
        A.chunk_base = 0;
        A.unperm = d_unperm;
        A.hits = d_hits;
        const int64_t tiles = (n_bound + per_block - 1) / per_block;
        dim3 grid((unsigned)tiles, (unsigned)A.items_per_tile, 1);
        {
            Timed tm(ctx, CAT_PERMUTE);
            sb::walk_permute_kernel<false><<<grid, tile_threads, smem, ctx->stream>>>(A);
            ctx->stats.kernel_launches += 1;
            if (!d_count) ctx->stats.tests_walks += n_bound * (int64_t)n_perms;   // else the device counts (d_walks)
            SB_CUDA(ctx, cudaGetLastError());
        }
        Timed tm(ctx, CAT_REDUCE);
        sb::accumulate_hits_kernel<<<(unsigned)((n_bound + 255) / 256), 256, 0, ctx->stream>>>(
            d_hits, S, d_list, (int32_t)n_bound, d_count, base, n_perms, P, ppi, early_stop, d_rmin, d_r, d_n_done,
            d_list_out, d_count_out, d_count ? ctx->d_walks : nullptr);
        ctx->stats.kernel_launches += 1;
        SB_CUDA(ctx, cudaGetLastError());
        return SB_OK;
    };
    if (!early_stop) {
        for (int base = 0; base < P; base += n_launch) {
            rc = slice(base, std::min(n_launch, P - base), nullptr, S, nullptr, nullptr, nullptr);
            if (rc) return rc;
        }
        return SB_OK;
    }
    // Reference-rule mode: walk the permutations in growing slices and keep only the genes the sequential rule has
    // not stopped yet.  The rule cannot fire before permutation 31 (methods.py:1360: i >= 30) and most null genes
    // stop right there, so the first slice is 31 labellings; the survivors of the first two slices (the host reads
    // their number, which bounds every later grid) then run slice after slice without a host round trip: each
    // round's list length stays on the device (S_dev) and blocks past the end exit at once.
    const int n_rounds_max = P + 4;      // every round walks at least one labelling
    rc = ensure_scratch(ctx, 8, sizeof(int32_t) * (2 * (size_t)S + (size_t)n_rounds_max + 4));
    if (rc) return rc;
    int32_t *d_list[2] = {(int32_t *)ctx->d_scratch[8], (int32_t *)ctx->d_scratch[8] + S};
    int32_t *d_counters = (int32_t *)ctx->d_scratch[8] + 2 * S;     // one per round
    SB_CUDA(ctx, cudaMemsetAsync(d_counters, 0, sizeof(int32_t) * (size_t)n_rounds_max, ctx->stream));
    if (!ctx->h_pinned_counter) SB_CUDA(ctx, cudaMallocHost(&ctx->h_pinned_counter, sizeof(int32_t)));
    int64_t n_bound = S;
    const int32_t *cur = nullptr, *cur_count = nullptr;   // null = identity list of S slots
    int base = 0, round = 0;
    while (base < P && n_bound > 0) {
        int n_perms = round == 0 ? 31 : 33, one = 1;
        if (round >= 2) plan_launch((n_bound + per_block - 1) / per_block, label_cap, slots, 1, 1, &one, &n_perms);
        if (round >= 2 && getenv("SB_RULE_ROUND_LABELLINGS")) n_perms = std::max(1, atoi(getenv("SB_RULE_ROUND_LABELLINGS")));   // tuning probe
        n_perms = std::min(std::min(n_perms, label_cap), P - base);
        int32_t *out = d_list[round & 1];
        rc = slice(base, n_perms, cur, n_bound, cur_count, out, d_counters + round);
        if (rc) return rc;
        cur = out;
        cur_count = d_counters + round;
        if (round < 4 || round % 4 == 0) {   // the rounds that remove most genes, then every fourth: read the survivor
                                             // count, so that the following rounds are sized for the genes really left
            SB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned_counter, cur_count, sizeof(int32_t), cudaMemcpyDeviceToHost,
                                         ctx->stream));
            SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            n_bound = *ctx->h_pinned_counter;
            cur_count = nullptr;     // exact on the host from here
            if (round == 3 && n_bound > 0 && base + n_perms < P) {
                // the slots that are left run all the remaining rounds: pack their gene columns side by side once
                genesC_pad = (n_bound + 31) / 32 * 32;
                rc = ensure_scratch(ctx, 14, sizeof(uint32_t) * (size_t)s.W32p * (size_t)genesC_pad + sizeof(int32_t) * (size_t)S);
                if (rc) return rc;
                uint32_t *gc = (uint32_t *)ctx->d_scratch[14];
                int32_t *cs = (int32_t *)(gc + (size_t)s.W32p * (size_t)genesC_pad);
                dim3 grid((unsigned)((n_bound + 255) / 256), (unsigned)std::min(s.W32p, 32), 1);
                sb::compact_columns_kernel<<<grid, 256, 0, ctx->stream>>>(s.d_genesT, s.Gs, s.W32p, d_gene_idx, cur, (int)n_bound,
                                                                         genesC_pad, gc, cs);
                ctx->stats.kernel_launches += 1;
                SB_CUDA(ctx, cudaGetLastError());
                d_genesC = gc;
                d_col_of_slot = cs;
            }
        }
        base += n_perms;
        ++round;
    }
    return SB_OK;
}

// ---- device epilogue (epilogue.cuh) -----------------------------------------------
// d_p [n], d_counts [n][4] or d_keep [n] (one of them) -> d_order [n] (first *m entries: tested genes by ascending p,
// stable), d_bonf [n], d_bh [n].  n_tests <= 0: the number of tested genes.  Synchronises once (the count).
int launch_adjust(sb_ctx *ctx, const double *d_p, const int32_t *d_counts, const uint8_t *d_keep, int64_t n,
                  int64_t n_tests, int32_t *d_order, double *d_bonf, double *d_bh, int64_t *m_out)
{
    if (n <= 0) { if (m_out) *m_out = 0; return SB_OK; }
    if (n > 0x7fffffffLL) return fail(ctx, SB_ERR_ARG, "sb_adjust_pvalues: at most 2^31 - 1 rows");
    const int n_chunks = (int)((n + sb::SORT_CHUNK - 1) / sb::SORT_CHUNK);
    const int64_t n_blocks = (n + 1023) / 1024;
    const size_t b_keys = sizeof(uint64_t) * (size_t)n, b_idx = sizeof(int32_t) * (size_t)n;
    const size_t b_hist = sizeof(uint32_t) * (size_t)sb::RADIX * n_chunks;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    int rc = ensure_scratch(ctx, 11, 2 * up(b_keys) + 2 * up(b_idx) + up(b_hist) + up(b_keys) + up(sizeof(double) * n_blocks) + 256);
    if (rc) return rc;
    char *base = (char *)ctx->d_scratch[11];
    uint64_t *keys[2] = {(uint64_t *)base, (uint64_t *)(base + up(b_keys))};
    base += 2 * up(b_keys);
    int32_t *idx[2] = {(int32_t *)base, (int32_t *)(base + up(b_idx))};
    base += 2 * up(b_idx);
    uint32_t *hist = (uint32_t *)base;
    base += up(b_hist);
    double *vals = (double *)base;
    base += up(b_keys);
    double *block_min = (double *)base;
    base += up(sizeof(double) * n_blocks);
    unsigned long long *d_kept = (unsigned long long *)base;
    Timed tm(ctx, CAT_EPILOGUE);
    SB_CUDA(ctx, cudaMemsetAsync(d_kept, 0, sizeof(unsigned long long), ctx->stream));
    const unsigned g256 = (unsigned)((n + 255) / 256);
    sb::epi_keys_kernel<<<g256, 256, 0, ctx->stream>>>(d_p, d_counts, d_keep, n, keys[0], idx[0], d_kept);
    const unsigned gsort = (unsigned)((n_chunks + sb::SORT_WARPS - 1) / sb::SORT_WARPS);
    int cur = 0;
    for (int shift = 0; shift < 64; shift += sb::RADIX_BITS) {
        sb::radix_hist_kernel<<<gsort, 32 * sb::SORT_WARPS, 0, ctx->stream>>>(keys[cur], n, shift, n_chunks, hist);
        sb::radix_scan_kernel<<<1, 1024, 0, ctx->stream>>>(hist, (int64_t)sb::RADIX * n_chunks);
        sb::radix_scatter_kernel<<<gsort, 32 * sb::SORT_WARPS, 0, ctx->stream>>>(keys[cur], idx[cur], n, shift, n_chunks, hist,
                                                                                keys[cur ^ 1], idx[cur ^ 1]);
        cur ^= 1;
    }
    ctx->stats.kernel_launches += 1 + 3 * (64 / sb::RADIX_BITS);
    SB_CUDA(ctx, cudaGetLastError());
    unsigned long long kept = 0;
    SB_CUDA(ctx, cudaMemcpyAsync(&kept, d_kept, sizeof kept, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int64_t m = (int64_t)kept;
    const double tests = (double)(n_tests > 0 ? n_tests : m);
    if (m > 0) {
        const int64_t mb = (m + 1023) / 1024;
        sb::bh_values_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(keys[cur], m, tests, vals);
        sb::suffix_min_block_kernel<<<(unsigned)mb, 1024, 0, ctx->stream>>>(vals, m, block_min);
        sb::suffix_min_top_kernel<<<1, 1024, 0, ctx->stream>>>(block_min, mb);
        ctx->stats.kernel_launches += 3;
    }
    sb::bh_finish_kernel<<<g256, 256, 0, ctx->stream>>>(keys[cur], idx[cur], n, m, tests, vals, block_min, d_order, d_bonf, d_bh);
    ctx->stats.kernel_launches += 1;
    SB_CUDA(ctx, cudaGetLastError());
    if (m_out) *m_out = m;
    return SB_OK;
}

int check_walk_ready(sb_ctx *ctx, int32_t t, int64_t S)
{
    if (t < 0 || t >= SB_MAX_TRAITS) return fail(ctx, SB_ERR_ARG, "trait index out of range");
    if (S <= 0) return fail(ctx, SB_ERR_ARG, "S must be positive");
    return finalize_slot(ctx, t);
}

}  // namespace

// ================================================================== C-ABI
extern "C" {

int sb_version(void) { return SB_VERSION; }

const char *sb_last_error(const sb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int sb_create(int device, sb_ctx **out)
{
    if (!out) { g_create_error = "sb_create: out is NULL"; return SB_ERR_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device available (libscoary_b200 has no CPU fallback): ") +
                         cudaGetErrorString(e);
        return SB_ERR_CUDA;
    }
    if (device < 0 || device >= n) { g_create_error = "sb_create: device ordinal out of range"; return SB_ERR_ARG; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return SB_ERR_CUDA; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return SB_ERR_CUDA; }
    if (prop.major != 10) {
        g_create_error = "libscoary_b200 is built for sm_100a (B200) only; found compute capability " +
                         std::to_string(prop.major) + "." + std::to_string(prop.minor);
        return SB_ERR_CUDA;
    }
    sb_ctx *ctx = new sb_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    ctx->stats.sm_count = ctx->sm_count;
    e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return SB_ERR_CUDA; }
    e = cudaMalloc(&ctx->d_walks, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_walks, 0, sizeof(unsigned long long));
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); cudaStreamDestroy(ctx->own_stream); delete ctx; return SB_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return SB_OK;
}

void sb_destroy(sb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    resolve_events(ctx);
    for (auto &e : ctx->free_events) { cudaEventDestroy(e.start); cudaEventDestroy(e.stop); }
    for (auto &s : ctx->traits) { free_trait(s); free_tree_storage(s); }
    cudaFree(ctx->d_genes_buf);
    cudaFree(ctx->d_trait_arena);
    cudaFree(ctx->d_lut);
    cudaFree(ctx->d_peak_out);
    cudaFree(ctx->d_walks);
    for (auto p : ctx->d_scratch) cudaFree(p);
    if (ctx->h_pinned_counter) cudaFreeHost(ctx->h_pinned_counter);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int sb_set_stream(sb_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return SB_OK;
}

int sb_synchronize(sb_ctx *ctx)
{
    if (!ctx) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

int sb_set_permute_mode(sb_ctx *ctx, int mode)
{
    if (!ctx || mode < 0 || mode > 2) return SB_ERR_ARG;
    ctx->permute_mode = mode;
    return SB_OK;
}

int sb_set_profiling(sb_ctx *ctx, int on)
{
    if (!ctx) return SB_ERR_ARG;
    ctx->profiling = on != 0;
    return SB_OK;
}

int sb_stats(sb_ctx *ctx, sb_stats_t *out)
{
    if (!ctx || !out) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    resolve_events(ctx);
    unsigned long long dev_walks = 0;
    SB_CUDA(ctx, cudaMemcpy(&dev_walks, ctx->d_walks, sizeof dev_walks, cudaMemcpyDeviceToHost));
    *out = ctx->stats;
    out->tests_walks += (int64_t)dev_walks;
    return SB_OK;
}

int sb_stats_reset(sb_ctx *ctx)
{
    if (!ctx) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    resolve_events(ctx);
    memset(&ctx->stats, 0, sizeof ctx->stats);
    ctx->stats.sm_count = ctx->sm_count;
    SB_CUDA(ctx, cudaMemset(ctx->d_walks, 0, sizeof(unsigned long long)));
    return SB_OK;
}

int sb_set_genes(sb_ctx *ctx, const uint64_t *bits, int64_t G, int32_t N, int32_t W)
{
    if (!ctx || !bits) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = set_genes_common(ctx, G, N, W);
    if (rc) return rc;
    const size_t bytes = sizeof(uint64_t) * (size_t)G * (size_t)W;
    if (bytes > ctx->genes_cap) {       // grow the context's own buffer; otherwise reuse it
        if (ctx->d_genes_buf) SB_CUDA(ctx, cudaFree(ctx->d_genes_buf));
        ctx->d_genes_buf = nullptr; ctx->genes_cap = 0;
        SB_CUDA(ctx, cudaMalloc(&ctx->d_genes_buf, bytes));
        ctx->genes_cap = bytes;
    }
    ctx->d_genes = ctx->d_genes_buf;
    ctx->own_genes = true;
    SB_CUDA(ctx, cudaMemcpyAsync(ctx->d_genes, bits, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += (int64_t)bytes;
    return SB_OK;
}

int sb_set_genes_device(sb_ctx *ctx, const uint64_t *d_bits, int64_t G, int32_t N, int32_t W)
{
    if (!ctx || !d_bits) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (((uintptr_t)d_bits) & 15) return fail(ctx, SB_ERR_ARG, "sb_set_genes_device: pointer must be 16-byte aligned");
    int rc = set_genes_common(ctx, G, N, W);
    if (rc) return rc;
    ctx->d_genes = const_cast<uint64_t *>(d_bits);
    ctx->own_genes = false;
    return SB_OK;
}

int sb_set_trait(sb_ctx *ctx, int32_t t, const uint64_t *value, const uint64_t *mask)
{
    if (!ctx || !value || !mask) return SB_ERR_ARG;
    if (t < 0 || t >= SB_MAX_TRAITS) return fail(ctx, SB_ERR_ARG, "trait index out of range");
    if (ctx->W == 0) return fail(ctx, SB_ERR_STATE, "call sb_set_genes before sb_set_trait");
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    TraitSlot &s = ctx->traits[t];
    const int W = ctx->W;
    s.h_value.assign(value, value + W);
    s.h_mask.assign(mask, mask + W);
    for (int w = 0; w < W; ++w) s.h_value[w] &= s.h_mask[w];
    const int tail = ctx->N & 63;
    for (int w = 0; w < W; ++w) {   // no bits beyond N
        uint64_t keep = ~0ULL;
        if (w > (ctx->N - 1) / 64) keep = 0;
        else if (w == (ctx->N - 1) / 64 && tail) keep = (1ULL << tail) - 1;
        s.h_value[w] &= keep;
        s.h_mask[w] &= keep;
    }
    s.d_value = ctx->d_trait_arena + (size_t)t * 2 * W;
    s.d_mask = s.d_value + W;
    SB_CUDA(ctx, cudaMemcpyAsync(s.d_value, s.h_value.data(), sizeof(uint64_t) * W, cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(s.d_mask, s.h_mask.data(), sizeof(uint64_t) * W, cudaMemcpyHostToDevice, ctx->stream));
    // no synchronisation: copies out of pageable memory return once the source has been staged, and later work on the
    // stream is ordered behind them
    ctx->stats.h2d_bytes += (int64_t)sizeof(uint64_t) * 2 * W;
    s.has_trait = true;
    s.finalized = false;   // labels depend on the trait; genesT does not
    return SB_OK;
}

int sb_set_tree(sb_ctx *ctx, int32_t t, const int32_t *left, const int32_t *right, int32_t n_internal,
                const int32_t *leaf_to_col)
{
    if (!ctx || !left || !right || !leaf_to_col) return SB_ERR_ARG;
    if (t < 0 || t >= SB_MAX_TRAITS) return fail(ctx, SB_ERR_ARG, "trait index out of range");
    if (n_internal < 1) return fail(ctx, SB_ERR_ARG, "sb_set_tree: need at least two leaves");
    if (n_internal + 1 > 32766) return fail(ctx, SB_ERR_ARG, "sb_set_tree: at most 32766 leaves (32-bit DP keys)");
    if (ctx->W == 0) return fail(ctx, SB_ERR_STATE, "call sb_set_genes before sb_set_tree");
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    Program prog;
    std::string err;
    if (!compile_tree(left, right, n_internal, prog, err)) return fail(ctx, SB_ERR_ARG, err.c_str());
    {   // the program and at least one label vector must fit the constant pool (walk.cuh): reject the tree HERE,
        // not at the first walk.  Every binary tree of <= 32 766 leaves passes (the longest program, a balanced
        // tree's 21 374 ops, leaves room for 5 vectors); the check guards builds with a padded leaf stream.
        const int64_t n_pos = (int64_t)prog.leaf_of_pos.size();
        const int64_t w32p = ((n_pos + 31) / 32 + 3) / 4 * 4;
        const int64_t need = sb::walk_label_base((int)prog.ops.size()) + w32p * std::max(1, sb::WALK_NLAB);
        if (need > sb::C_POOL_WORDS) {
            char buf[200];
            snprintf(buf, sizeof buf, "sb_set_tree: tree program (%d ops) and one label vector (%lld words) exceed the "
                     "%d-word constant pool", (int)prog.ops.size(), (long long)w32p, sb::C_POOL_WORDS);
            return fail(ctx, SB_ERR_ARG, buf);
        }
    }
    TraitSlot &s = ctx->traits[t];
    free_tree(s);
    s.n_internal = n_internal;
    s.n_leaves = n_internal + 1;
    s.W32 = (s.n_leaves + 31) / 32;
    s.W32p = (s.W32 + 3) / 4 * 4;
    s.depth = prog.depth;
    s.shift = 1;
    while ((1 << s.shift) <= s.n_leaves / 2) ++s.shift;   // 2^SH > max pairs (= floor(n/2)) >= pro, anti
    s.n_ops = (int32_t)prog.ops.size();
    s.lab_base = sb::walk_label_base(s.n_ops);
    s.h_leaf_to_col.assign(leaf_to_col, leaf_to_col + s.n_leaves);
    // stream positions: the leaves in walk order; with SB_WALK_PADDED also pad positions (-1), and the label /
    // gene words in walk order (W32p) are counted over positions.  The device copies are filled up with -1 to
    // W32p * 32 entries so that the kernels need no separate position count.
    int32_t n_pos = (int32_t)prog.leaf_of_pos.size();
    if (sb::WALK_PADDED) s.W32p = ((n_pos + 31) / 32 + 3) / 4 * 4;
    const int32_t n_alloc = sb::WALK_PADDED ? s.W32p * 32 : s.n_leaves;
    s.h_leaf_of_pos = prog.leaf_of_pos;
    std::vector<int32_t> walk_col((size_t)n_alloc, -1), pos_leaf((size_t)n_alloc, -1);
    for (int32_t pos = 0; pos < n_pos; ++pos) {
        const int32_t leaf = prog.leaf_of_pos[pos];
        if (leaf < 0) continue;      // pad position
        const int32_t col = leaf_to_col[leaf];
        if (col < 0 || col >= ctx->N) return fail(ctx, SB_ERR_ARG, "sb_set_tree: leaf_to_col out of range");
        walk_col[pos] = col;
        pos_leaf[pos] = leaf;
    }
    s.h_ops = prog.ops;
    SB_CUDA(ctx, cudaMalloc(&s.d_walk_col, sizeof(int32_t) * n_alloc));
    SB_CUDA(ctx, cudaMalloc(&s.d_leaf_of_pos, sizeof(int32_t) * n_alloc));
    SB_CUDA(ctx, cudaMemcpyAsync(s.d_walk_col, walk_col.data(), sizeof(int32_t) * n_alloc, cudaMemcpyHostToDevice,
                                 ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(s.d_leaf_of_pos, pos_leaf.data(), sizeof(int32_t) * n_alloc,
                                 cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.h2d_bytes += (int64_t)(2 * sizeof(int32_t) * n_alloc);
    s.has_tree = true;
    s.finalized = false;
    return SB_OK;
}

int sb_contingency_fisher_device(sb_ctx *ctx, int32_t t, int32_t *d_counts, double *d_p, uint64_t *d_hash)
{
    if (!ctx) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_fisher(ctx, t, 1, d_counts, d_p, d_hash);
}

int sb_contingency_fisher_multi_device(sb_ctx *ctx, int32_t t0, int32_t n_traits, int32_t *d_counts, double *d_p,
                                       uint64_t *d_hash)
{
    if (!ctx) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_fisher(ctx, t0, n_traits, d_counts, d_p, d_hash);
}

int sb_contingency_fisher_multi(sb_ctx *ctx, int32_t t0, int32_t n_traits, int32_t *counts, double *p, uint64_t *hash)
{
    if (!ctx) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t G = (size_t)ctx->G;
    if (G == 0) return fail(ctx, SB_ERR_STATE, "genes not set (sb_set_genes)");
    if (n_traits < 1 || t0 < 0 || t0 + n_traits > SB_MAX_TRAITS) return fail(ctx, SB_ERR_ARG, "trait index out of range");
    const size_t GT = G * (size_t)n_traits;
    int rc = ensure_scratch(ctx, 2, GT * (16 + 8 + 16));
    if (rc) return rc;
    char *base = (char *)ctx->d_scratch[2];
    int32_t *d_counts = (int32_t *)base;
    double *d_p = (double *)(base + GT * 16);
    uint64_t *d_hash = (uint64_t *)(base + GT * 24);
    rc = launch_fisher(ctx, t0, n_traits, d_counts, p ? d_p : nullptr, hash ? d_hash : nullptr);
    if (rc) return rc;
    if (counts) SB_CUDA(ctx, cudaMemcpyAsync(counts, d_counts, GT * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (p) SB_CUDA(ctx, cudaMemcpyAsync(p, d_p, GT * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (hash) SB_CUDA(ctx, cudaMemcpyAsync(hash, d_hash, GT * 16, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += (int64_t)(GT * ((counts ? 16 : 0) + (p ? 8 : 0) + (hash ? 16 : 0)));
    return SB_OK;
}

int sb_contingency_fisher(sb_ctx *ctx, int32_t t, int32_t *counts, double *p, uint64_t *hash)
{
    return sb_contingency_fisher_multi(ctx, t, 1, counts, p, hash);
}

int sb_pairwise_device(sb_ctx *ctx, int32_t t, const int64_t *d_gene_idx, int64_t S, int32_t *d_pairs)
{
    if (!ctx || !d_pairs) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = check_walk_ready(ctx, t, S);
    if (rc) return rc;
    if (!d_gene_idx && S > ctx->G) return fail(ctx, SB_ERR_ARG, "S exceeds the number of genes");
    return launch_pairwise(ctx, t, d_gene_idx, S, d_pairs);
}

static int upload_gene_idx(sb_ctx *ctx, const int64_t *gene_idx, int64_t S, const int64_t **d_out)
{
    *d_out = nullptr;
    if (!gene_idx) {
        if (S > ctx->G) return fail(ctx, SB_ERR_ARG, "S exceeds the number of genes");
        return SB_OK;
    }
    for (int64_t i = 0; i < S; ++i)
        if (gene_idx[i] < 0 || gene_idx[i] >= ctx->G) return fail(ctx, SB_ERR_ARG, "gene_idx out of range");
    int rc = ensure_scratch(ctx, 3, sizeof(int64_t) * (size_t)S);
    if (rc) return rc;
    SB_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch[3], gene_idx, sizeof(int64_t) * (size_t)S, cudaMemcpyHostToDevice,
                                 ctx->stream));
    ctx->stats.h2d_bytes += (int64_t)sizeof(int64_t) * S;
    *d_out = (const int64_t *)ctx->d_scratch[3];
    return SB_OK;
}

int sb_pairwise(sb_ctx *ctx, int32_t t, const int64_t *gene_idx, int64_t S, int32_t *pairs)
{
    if (!ctx || !pairs) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = check_walk_ready(ctx, t, S);
    if (rc) return rc;
    const int64_t *d_idx;
    rc = upload_gene_idx(ctx, gene_idx, S, &d_idx);
    if (rc) return rc;
    rc = ensure_scratch(ctx, 4, sizeof(int32_t) * 3 * (size_t)S);
    if (rc) return rc;
    int32_t *d_pairs = (int32_t *)ctx->d_scratch[4];
    rc = launch_pairwise(ctx, t, d_idx, S, d_pairs);
    if (rc) return rc;
    SB_CUDA(ctx, cudaMemcpyAsync(pairs, d_pairs, sizeof(int32_t) * 3 * (size_t)S, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += (int64_t)sizeof(int32_t) * 3 * S;
    return SB_OK;
}

int sb_permute_device(sb_ctx *ctx, int32_t t, const int64_t *d_gene_idx, int64_t S, int32_t P, uint64_t seed,
                      int32_t early_stop, const int32_t *d_rmin, int32_t *d_pairs, int32_t *d_r, int32_t *d_n_done)
{
    if (!ctx || !d_r || !d_n_done) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (P < 1) return fail(ctx, SB_ERR_ARG, "P must be >= 1");
    if (early_stop && !d_rmin) return fail(ctx, SB_ERR_ARG, "early_stop needs rmin");
    int rc = check_walk_ready(ctx, t, S);
    if (rc) return rc;
    if (!d_gene_idx && S > ctx->G) return fail(ctx, SB_ERR_ARG, "S exceeds the number of genes");
    int32_t *d_un = d_pairs;
    if (!d_un) {
        rc = ensure_scratch(ctx, 4, sizeof(int32_t) * 3 * (size_t)S);
        if (rc) return rc;
        d_un = (int32_t *)ctx->d_scratch[4];
    }
    rc = launch_pairwise(ctx, t, d_gene_idx, S, d_un);
    if (rc) return rc;
    return launch_permute(ctx, t, d_gene_idx, S, P, seed, early_stop, d_rmin, d_un, d_r, d_n_done);
}

int sb_permute_range_device(sb_ctx *ctx, int32_t t, const int64_t *d_gene_idx, int64_t S, int32_t perm_first,
                            int32_t perm_count, uint64_t seed, int32_t *d_pairs, int32_t *d_r)
{
    if (!ctx || !d_r) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (perm_count < 1 || perm_first < 0) return fail(ctx, SB_ERR_ARG, "sb_permute_range: need perm_first >= 0 and perm_count >= 1");
    int rc = check_walk_ready(ctx, t, S);
    if (rc) return rc;
    if (!d_gene_idx && S > ctx->G) return fail(ctx, SB_ERR_ARG, "S exceeds the number of genes");
    int32_t *d_un = d_pairs;
    if (!d_un) {
        rc = ensure_scratch(ctx, 4, sizeof(int32_t) * 3 * (size_t)S);
        if (rc) return rc;
        d_un = (int32_t *)ctx->d_scratch[4];
    }
    rc = ensure_scratch(ctx, 5, sizeof(int32_t) * 2 * (size_t)S);
    if (rc) return rc;
    rc = launch_pairwise(ctx, t, d_gene_idx, S, d_un);
    if (rc) return rc;
    return launch_permute(ctx, t, d_gene_idx, S, perm_count, seed, 0, nullptr, d_un, d_r, (int32_t *)ctx->d_scratch[5] + S,
                          perm_first);
}

int sb_permute_range(sb_ctx *ctx, int32_t t, const int64_t *gene_idx, int64_t S, int32_t perm_first, int32_t perm_count,
                     uint64_t seed, int32_t *pairs, int32_t *r)
{
    if (!ctx || !r) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (perm_count < 1 || perm_first < 0) return fail(ctx, SB_ERR_ARG, "sb_permute_range: need perm_first >= 0 and perm_count >= 1");
    int rc = check_walk_ready(ctx, t, S);
    if (rc) return rc;
    const int64_t *d_idx;
    rc = upload_gene_idx(ctx, gene_idx, S, &d_idx);
    if (rc) return rc;
    rc = ensure_scratch(ctx, 4, sizeof(int32_t) * 3 * (size_t)S);
    if (rc) return rc;
    rc = ensure_scratch(ctx, 5, sizeof(int32_t) * 2 * (size_t)S);
    if (rc) return rc;
    int32_t *d_un = (int32_t *)ctx->d_scratch[4];
    int32_t *d_r = (int32_t *)ctx->d_scratch[5];
    rc = launch_pairwise(ctx, t, d_idx, S, d_un);
    if (rc) return rc;
    rc = launch_permute(ctx, t, d_idx, S, perm_count, seed, 0, nullptr, d_un, d_r, d_r + S, perm_first);
    if (rc) return rc;
    if (pairs) SB_CUDA(ctx, cudaMemcpyAsync(pairs, d_un, sizeof(int32_t) * 3 * (size_t)S, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(r, d_r, sizeof(int32_t) * (size_t)S, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += (int64_t)sizeof(int32_t) * S * (pairs ? 4 : 1);
    return SB_OK;
}

int sb_permute(sb_ctx *ctx, int32_t t, const int64_t *gene_idx, int64_t S, int32_t P, uint64_t seed,
               int32_t early_stop, const int32_t *rmin, int32_t *pairs, int32_t *r, int32_t *n_done)
{
    if (!ctx || !r || !n_done) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (P < 1) return fail(ctx, SB_ERR_ARG, "P must be >= 1");
    if (early_stop && !rmin) return fail(ctx, SB_ERR_ARG, "early_stop needs rmin");
    int rc = check_walk_ready(ctx, t, S);
    if (rc) return rc;
    const int64_t *d_idx;
    rc = upload_gene_idx(ctx, gene_idx, S, &d_idx);
    if (rc) return rc;
    // scratch 4: unperm [S][3]; scratch 5: r [S], n_done [S]; scratch 6: rmin [P]
    rc = ensure_scratch(ctx, 4, sizeof(int32_t) * 3 * (size_t)S);
    if (rc) return rc;
    rc = ensure_scratch(ctx, 5, sizeof(int32_t) * 2 * (size_t)S);
    if (rc) return rc;
    int32_t *d_un = (int32_t *)ctx->d_scratch[4];
    int32_t *d_r = (int32_t *)ctx->d_scratch[5];
    int32_t *d_nd = d_r + S;
    const int32_t *d_rmin = nullptr;
    if (early_stop) {
        rc = ensure_scratch(ctx, 6, sizeof(int32_t) * (size_t)P);
        if (rc) return rc;
        SB_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch[6], rmin, sizeof(int32_t) * (size_t)P, cudaMemcpyHostToDevice,
                                     ctx->stream));
        ctx->stats.h2d_bytes += (int64_t)sizeof(int32_t) * P;
        d_rmin = (const int32_t *)ctx->d_scratch[6];
    }
    rc = launch_pairwise(ctx, t, d_idx, S, d_un);
    if (rc) return rc;
    rc = launch_permute(ctx, t, d_idx, S, P, seed, early_stop, d_rmin, d_un, d_r, d_nd);
    if (rc) return rc;
    if (pairs) SB_CUDA(ctx, cudaMemcpyAsync(pairs, d_un, sizeof(int32_t) * 3 * (size_t)S, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(r, d_r, sizeof(int32_t) * (size_t)S, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(n_done, d_nd, sizeof(int32_t) * (size_t)S, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += (int64_t)sizeof(int32_t) * S * (pairs ? 5 : 2);
    return SB_OK;
}

int sb_adjust_pvalues_device(sb_ctx *ctx, const double *d_p, const int32_t *d_counts, int64_t n, int64_t n_tests,
                             int32_t *d_order, double *d_bonferroni, double *d_bh, int64_t *n_tested)
{
    if (!ctx || !d_p || !d_counts) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_adjust(ctx, d_p, d_counts, nullptr, n, n_tests, d_order, d_bonferroni, d_bh, n_tested);
}

int sb_adjust_pvalues(sb_ctx *ctx, const double *p, const uint8_t *keep, int64_t n, int64_t n_tests, int32_t *order,
                      double *bonferroni, double *bh, int64_t *n_tested)
{
    if (!ctx || !p || !keep) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n <= 0) { if (n_tested) *n_tested = 0; return SB_OK; }
    const size_t N = (size_t)n;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    int rc = ensure_scratch(ctx, 12, 3 * up(8 * N) + up(4 * N) + up(N));
    if (rc) return rc;
    char *base = (char *)ctx->d_scratch[12];
    double *d_p = (double *)base, *d_bonf = (double *)(base + up(8 * N)), *d_bh = (double *)(base + 2 * up(8 * N));
    int32_t *d_order = (int32_t *)(base + 3 * up(8 * N));
    uint8_t *d_keep = (uint8_t *)(base + 3 * up(8 * N) + up(4 * N));
    SB_CUDA(ctx, cudaMemcpyAsync(d_p, p, 8 * N, cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(d_keep, keep, N, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += (int64_t)(9 * N);
    rc = launch_adjust(ctx, d_p, nullptr, d_keep, n, n_tests, d_order, d_bonf, d_bh, n_tested);
    if (rc) return rc;
    if (order) SB_CUDA(ctx, cudaMemcpyAsync(order, d_order, 4 * N, cudaMemcpyDeviceToHost, ctx->stream));
    if (bonferroni) SB_CUDA(ctx, cudaMemcpyAsync(bonferroni, d_bonf, 8 * N, cudaMemcpyDeviceToHost, ctx->stream));
    if (bh) SB_CUDA(ctx, cudaMemcpyAsync(bh, d_bh, 8 * N, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += (int64_t)((order ? 4 : 0) + (bonferroni ? 8 : 0) + (bh ? 8 : 0)) * n;
    return SB_OK;
}

int sb_binom_two_sided(sb_ctx *ctx, const int32_t *k, const int32_t *n, int64_t count, double *p)
{
    if (!ctx || !k || !n || !p) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count <= 0) return SB_OK;
    int32_t n_max = 0;
    for (int64_t i = 0; i < count; ++i) n_max = std::max(n_max, n[i]);
    int rc = ensure_lut(ctx, n_max);
    if (rc) return rc;
    const size_t C = (size_t)count;
    rc = ensure_scratch(ctx, 13, 16 * C + 512);
    if (rc) return rc;
    int32_t *d_k = (int32_t *)ctx->d_scratch[13], *d_n = d_k + C;
    double *d_out = (double *)((char *)ctx->d_scratch[13] + (8 * C + 255) / 256 * 256);
    SB_CUDA(ctx, cudaMemcpyAsync(d_k, k, 4 * C, cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(d_n, n, 4 * C, cudaMemcpyHostToDevice, ctx->stream));
    {
        Timed tm(ctx, CAT_EPILOGUE);
        sb::binom_two_sided_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(d_k, d_n, 1, count, ctx->d_lut, d_out);
        ctx->stats.kernel_launches += 1;
        SB_CUDA(ctx, cudaGetLastError());
    }
    SB_CUDA(ctx, cudaMemcpyAsync(p, d_out, 8 * C, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.h2d_bytes += (int64_t)(8 * C);
    ctx->stats.d2h_bytes += (int64_t)(8 * C);
    return SB_OK;
}

int sb_debug_shuffled_labels(sb_ctx *ctx, int32_t t, int32_t P, uint64_t seed, uint8_t *labels)
{
    if (!ctx || !labels) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (P < 1) return fail(ctx, SB_ERR_ARG, "P must be >= 1");
    int rc = check_walk_ready(ctx, t, 1);
    if (rc) return rc;
    TraitSlot &s = ctx->traits[t];
    rc = ensure_scratch(ctx, 0, sizeof(uint32_t) * (size_t)P * s.W32p);
    if (rc) return rc;
    rc = ensure_scratch(ctx, 7, (size_t)P * s.n_leaves);
    if (rc) return rc;
    rc = launch_shuffle(ctx, t, P, 0, seed, (uint32_t *)ctx->d_scratch[0], (uint8_t *)ctx->d_scratch[7]);
    if (rc) return rc;
    SB_CUDA(ctx, cudaMemcpyAsync(labels, ctx->d_scratch[7], (size_t)P * s.n_leaves, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

int sb_upgma(sb_ctx *ctx, int32_t *merges)
{
    if (!ctx || !merges) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->d_genes) return fail(ctx, SB_ERR_STATE, "genes not set (sb_set_genes)");
    const int N = ctx->N, W = ctx->W;
    const int64_t G = ctx->G;
    if (N < 2) return fail(ctx, SB_ERR_ARG, "need at least two isolates to build a tree");
    if (N > 32766) return fail(ctx, SB_ERR_ARG, "sb_upgma: at most 32766 isolates");
    const int64_t Gw = (G + 63) / 64;
    // device buffers (freed on exit; tree building runs once per data set)
    uint64_t *d_T = nullptr, *d_allowed = nullptr;
    unsigned long long *d_nvar = nullptr;
    sb::UpgmaState S;
    memset(&S, 0, sizeof S);
    S.N = N;
    std::vector<uint64_t> h_allowed(W, 0);
    for (int j = 0; j < N; ++j) h_allowed[j >> 6] |= 1ULL << (j & 63);
    auto cleanup = [&]() {
        cudaFree(d_T); cudaFree(d_allowed); cudaFree(d_nvar); cudaFree(S.D); cudaFree(S.rowmin_val); cudaFree(S.rowmin_col);
        cudaFree(S.size); cudaFree(S.alive); cudaFree(S.redo); cudaFree(S.pick); cudaFree(S.merges);
    };
#define SB_TRY(call)                                                \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) {                                   \
            cleanup();                                              \
            ctx->err = std::string(#call) + " failed: " + cudaGetErrorString(e__); \
            return SB_ERR_CUDA;                                     \
        }                                                           \
    } while (0)
    SB_TRY(cudaMalloc(&d_T, sizeof(uint64_t) * (size_t)N * (size_t)Gw));
    SB_TRY(cudaMalloc(&d_allowed, sizeof(uint64_t) * W));
    SB_TRY(cudaMalloc(&d_nvar, sizeof(unsigned long long)));
    SB_TRY(cudaMalloc(&S.D, sizeof(double) * (size_t)N * (size_t)N));
    SB_TRY(cudaMalloc(&S.rowmin_val, sizeof(double) * N));
    SB_TRY(cudaMalloc(&S.rowmin_col, sizeof(int) * N));
    SB_TRY(cudaMalloc(&S.size, sizeof(double) * N));
    SB_TRY(cudaMalloc(&S.alive, N));
    SB_TRY(cudaMalloc(&S.redo, N));
    SB_TRY(cudaMalloc(&S.pick, sizeof(int) * 4));
    SB_TRY(cudaMemsetAsync(S.pick, 0, sizeof(int) * 4, ctx->stream));
    SB_TRY(cudaMalloc(&S.merges, sizeof(int) * 2 * (size_t)(N - 1)));
    SB_TRY(cudaMemcpyAsync(d_allowed, h_allowed.data(), sizeof(uint64_t) * W, cudaMemcpyHostToDevice, ctx->stream));
    SB_TRY(cudaMemsetAsync(d_nvar, 0, sizeof(unsigned long long), ctx->stream));
    std::vector<double> ones(N, 1.0);
    SB_TRY(cudaMemcpyAsync(S.size, ones.data(), sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
    SB_TRY(cudaMemsetAsync(S.alive, 1, N, ctx->stream));
    {
        Timed tm(ctx, CAT_TREE);
        dim3 g1((unsigned)Gw, (unsigned)((W + 3) / 4), 1);
        sb::transpose_variable_kernel<<<g1, 256, 0, ctx->stream>>>(ctx->d_genes, G, W, N, d_allowed, N, Gw, d_T, d_nvar);
        const unsigned tiles = (unsigned)((N + 31) / 32);
        sb::hamming_kernel<<<dim3(tiles, tiles, 1), dim3(32, 32, 1), 0, ctx->stream>>>(d_T, N, Gw, d_nvar, d_allowed, S.D);
        sb::upgma_rowmin_all_kernel<<<N, 256, 0, ctx->stream>>>(S);
        ctx->stats.kernel_launches += 3;
        // The N - 1 merge steps are four small kernels each (~20 000 launches at N = 5 000): one CUDA graph of
        // UPGMA_GRAPH_STEPS steps is captured once and replayed; the kernels read the step counter from device
        // memory and the steps past N - 1 in the last replay do nothing.
        auto one_step = [&]() {
            sb::upgma_pick_kernel<<<1, 1024, 0, ctx->stream>>>(S);
            sb::upgma_update_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(S);
            sb::upgma_redo_kernel<<<N, 256, 0, ctx->stream>>>(S);
            sb::upgma_finish_step_kernel<<<1, 1, 0, ctx->stream>>>(S);
        };
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        bool graphed = false;
        if (N - 1 > sb::UPGMA_GRAPH_STEPS &&
            cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            for (int k = 0; k < sb::UPGMA_GRAPH_STEPS; ++k) one_step();
            graphed = cudaStreamEndCapture(ctx->stream, &graph) == cudaSuccess && graph &&
                      cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
        }
        if (graphed) {
            for (int step = 0; step < N - 1; step += sb::UPGMA_GRAPH_STEPS) SB_TRY(cudaGraphLaunch(exec, ctx->stream));
        } else {
            (void)cudaGetLastError();
            for (int step = 0; step < N - 1; ++step) one_step();
        }
        SB_TRY(cudaStreamSynchronize(ctx->stream));
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        ctx->stats.kernel_launches += 4LL * (N - 1);
    }
    SB_TRY(cudaGetLastError());
    SB_TRY(cudaMemcpyAsync(merges, S.merges, sizeof(int) * 2 * (size_t)(N - 1), cudaMemcpyDeviceToHost, ctx->stream));
    SB_TRY(cudaStreamSynchronize(ctx->stream));
#undef SB_TRY
    ctx->stats.d2h_bytes += (int64_t)sizeof(int) * 2 * (N - 1);
    cleanup();
    return SB_OK;
}

// Host-only test hook (no GPU needed): compile a tree into the K4/K5 stack program.
// ops_out [max_ops] receives the uint16 program, leaf_of_pos [n_internal + 1] the walk order;
// returns the number of ops (> 0), or a negative sb_status; *stack_units = DP stack depth.
int sb_debug_compile_tree(const int32_t *left, const int32_t *right, int32_t n_internal, uint16_t *ops_out,
                          int32_t max_ops, int32_t *leaf_of_pos, int32_t *stack_units)
{
    if (!left || !right || !ops_out || !leaf_of_pos || n_internal < 1) return SB_ERR_ARG;
    Program prog;
    std::string err;
    if (!compile_tree(left, right, n_internal, prog, err)) { g_create_error = err; return SB_ERR_ARG; }
    if ((int32_t)prog.ops.size() > max_ops) { g_create_error = "program longer than max_ops"; return SB_ERR_ARG; }
    if ((int32_t)prog.leaf_of_pos.size() != n_internal + 1) {
        g_create_error = "this build pads the leaf stream: use sb_debug_compile_tree2";
        return SB_ERR_STATE;
    }
    memcpy(ops_out, prog.ops.data(), sizeof(uint16_t) * prog.ops.size());
    memcpy(leaf_of_pos, prog.leaf_of_pos.data(), sizeof(int32_t) * (size_t)(n_internal + 1));
    if (stack_units) *stack_units = prog.depth;
    return (int)prog.ops.size();
}

// As sb_debug_compile_tree, for builds whose stream may hold pad positions (SB_WALK_PADDED): leaf_of_pos
// [max_pos] receives the stream (leaf id, or -1 for a pad position) and *n_pos its length.
int sb_debug_compile_tree2(const int32_t *left, const int32_t *right, int32_t n_internal, uint16_t *ops_out,
                           int32_t max_ops, int32_t *leaf_of_pos, int32_t max_pos, int32_t *n_pos,
                           int32_t *stack_units)
{
    if (!left || !right || !ops_out || !leaf_of_pos || !n_pos || n_internal < 1) return SB_ERR_ARG;
    Program prog;
    std::string err;
    if (!compile_tree(left, right, n_internal, prog, err)) { g_create_error = err; return SB_ERR_ARG; }
    if ((int32_t)prog.ops.size() > max_ops) { g_create_error = "program longer than max_ops"; return SB_ERR_ARG; }
    if ((int32_t)prog.leaf_of_pos.size() > max_pos) { g_create_error = "stream longer than max_pos"; return SB_ERR_ARG; }
    memcpy(ops_out, prog.ops.data(), sizeof(uint16_t) * prog.ops.size());
    memcpy(leaf_of_pos, prog.leaf_of_pos.data(), sizeof(int32_t) * prog.leaf_of_pos.size());
    *n_pos = (int32_t)prog.leaf_of_pos.size();
    if (stack_units) *stack_units = prog.depth;
    return (int)prog.ops.size();
}

int sb_int32_peak(sb_ctx *ctx, int32_t iters, double *ops_per_s)
{
    if (!ctx || !ops_per_s || iters < 1) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->d_peak_out) SB_CUDA(ctx, cudaMalloc(&ctx->d_peak_out, sizeof(int)));
    const int blocks = ctx->sm_count * 8, threads = 256;
    cudaEvent_t e0, e1;
    SB_CUDA(ctx, cudaEventCreate(&e0));
    SB_CUDA(ctx, cudaEventCreate(&e1));
    sb::int32_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->d_peak_out, 16, 12345);   // warm-up
    SB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    sb::int32_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->d_peak_out, iters, 12345);
    SB_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    SB_CUDA(ctx, cudaEventSynchronize(e1));
    ctx->stats.kernel_launches += 2;
    float ms = 0.f;
    SB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // one VIADDMNMX = one add + one max; 64 of them per thread per iteration
    const double ops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    *ops_per_s = ops / ((double)ms * 1e-3);
    return SB_OK;
}


// undocumented design probe: warp-instructions per second for 8 instruction kinds
int sb_debug_pipe_rates(sb_ctx *ctx, int32_t iters, double *out8)
{
    if (!ctx || !out8 || iters < 1) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->d_peak_out) SB_CUDA(ctx, cudaMalloc(&ctx->d_peak_out, sizeof(int)));
    const int blocks = ctx->sm_count * 8, threads = 256;
    cudaEvent_t e0, e1;
    SB_CUDA(ctx, cudaEventCreate(&e0));
    SB_CUDA(ctx, cudaEventCreate(&e1));
    for (int mode = 0; mode < 8; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            if (rep) SB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
            const int it = rep ? iters : 16;
            switch (mode) {
                case 0: sb::pipe_rate_kernel<0><<<blocks, threads, 0, ctx->stream>>>(ctx->d_peak_out, it, 12345); break;
                case 1: sb::pipe_rate_kernel<1><<<blocks, threads, 0, ctx->stream>>>(ctx->d_peak_out, it, 12345); break;
                case 2: sb::pipe_rate_kernel<2><<<blocks, threads, 0, ctx->stream>>>(ctx->d_peak_out, it, 12345); break;
                case 3: sb::pipe_rate_kernel<3><<<blocks, threads, 0, ctx->stream>>>(ctx->d_peak_out, it, 12345); break;
                case 4: sb::pipe_rate_kernel<4><<<blocks, threads, 0, ctx->stream>>>(ctx->d_peak_out, it, 12345); break;
                case 5: sb::pipe_rate_kernel<5><<<blocks, threads, 0, ctx->stream>>>(ctx->d_peak_out, it, 12345); break;
                case 6: sb::pipe_rate_kernel<6><<<blocks, threads, 0, ctx->stream>>>(ctx->d_peak_out, it, 12345); break;
                default: sb::pipe_rate_kernel<7><<<blocks, threads, 0, ctx->stream>>>(ctx->d_peak_out, it, 12345); break;
            }
        }
        SB_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        SB_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        SB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        out8[mode] = 64.0 * (double)iters * (double)blocks * (double)threads / 32.0 / ((double)ms * 1e-3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return SB_OK;
}

}  // extern "C"
