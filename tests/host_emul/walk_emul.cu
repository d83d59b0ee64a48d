// TEST INFRASTRUCTURE: the walk kernels of scoary_b200/csrc/walk.cuh compiled for the HOST.
//
// walk.cuh is included with SB_HOST_EMUL: __device__ functions become plain inline functions, the
// __constant__ program / label arrays become host arrays, threadIdx / blockIdx become variables this
// driver sets, and the two kernels (walk_pairs_kernel, walk_permute_kernel) become functions that are
// called once per simulated thread.  Threads of a block only share the layout of the DP stack
// ([slot][field][thread]), never data, so running them one after the other is exact.  The DPX
// intrinsics (__viaddmax_s32, __vimax3_s16x2, ...) have host implementations in the CUDA headers;
// __vadd2 / __vmaxs2 do not and are spelled out in walk.cuh's emulation block.
//
// Built by tests/test_host_emul.py with `nvcc -DSB_HOST_EMUL -shared` (no device code is generated
// for these functions); nothing here is product code.
#define SB_HOST_EMUL 1
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../scoary_b200/csrc/walk.cuh"

namespace {

void fill(sb::WalkArgs &A, const uint32_t *genesT, int64_t Gs, int64_t S, int W32p, int shift)
{
    memset(&A, 0, sizeof A);
    A.genesT = genesT; A.Gs = Gs; A.gene_idx = nullptr; A.slot_idx = nullptr; A.S = S; A.S_total = S;
    A.W32p = W32p; A.shift = shift; A.tile_threads = sb::WALK_THREADS;
}

// the simulated shared memory is followed by guard words: a kernel that needs more stack than the compiler
// reported (stack_units) tramples them and the call fails
constexpr size_t GUARD = 4096;
constexpr int CANARY = 0x5CA7A11;

bool guard_intact(const std::vector<int> &smem, size_t words)
{
    for (size_t i = words; i < smem.size(); ++i)
        if (smem[i] != CANARY) return false;
    return true;
}

// program at the start of the constant pool, n_vec label vectors behind it (as engine.cu lays them out);
// returns the label base, or -1 if they do not fit
int load_pool(const uint16_t *ops, int n_ops, const uint32_t *labels, int64_t n_vec, int W32p)
{
    const int base = sb::walk_label_base(n_ops);
    if ((int64_t)base + n_vec * W32p > sb::C_POOL_WORDS) return -1;
    memcpy(sb::c_pool, ops, sizeof(uint16_t) * (size_t)n_ops);
    memcpy(sb::c_pool + base, labels, sizeof(uint32_t) * (size_t)n_vec * W32p);
    return base;
}

}  // namespace

extern "C" {

int emul_walk_threads(void) { return sb::WALK_THREADS; }
long long emul_carry_violations(void) { return sb::sb_emul_carry_violations; }
int emul_walk_genes_per_thread(void) { return sb::WALK_NP; }

// K4: pairs[S][3] for the labelling `labels` (walk order, W32p words)
int emul_pairs(const uint16_t *ops, int n_ops, const uint32_t *labels, const uint32_t *genesT, int64_t Gs, int64_t S,
               int W32p, int shift, int stack_units, int32_t *pairs)
{
    const int base = load_pool(ops, n_ops, labels, 1, W32p);
    if (base < 0) return -1;
    sb::WalkArgs A;
    fill(A, genesT, Gs, S, W32p, shift);
    A.pairs = pairs; A.lab_base = base;
    const int T = sb::WALK_THREADS;
    const size_t words = (size_t)sb::walk_push_words(true, 1) * (stack_units > 0 ? stack_units : 1) * T;
    std::vector<int> smem(words + GUARD, CANARY);
    const int64_t per_block = (int64_t)T * sb::WALK_NP;
    for (int64_t tile = 0; tile < (S + per_block - 1) / per_block; ++tile)
        for (int tid = 0; tid < T; ++tid) {
            sb::blockIdx = {(int)tile, 0, 0};
            sb::threadIdx = {tid, 0, 0};
            sb::sb_emul_shared = smem.data();
            sb::walk_pairs_kernel(A);
        }
    return guard_intact(smem, words) ? 0 : -2;
}

// K5: hits[ceil(P / ppi)][S] for the labellings labelsW[P][W32p] (walk order)
int emul_permute(const uint16_t *ops, int n_ops, const uint32_t *labelsW, int P, int ppi, const uint32_t *genesT,
                 int64_t Gs, int64_t S, int W32p, int shift, int stack_units, const int32_t *unperm, uint8_t *hits)
{
    const int base = load_pool(ops, n_ops, labelsW, P, W32p);
    if (base < 0 || ppi < 1 || ppi > sb::PERMS_PER_ITEM_MAX) return -1;
    sb::WalkArgs A;
    fill(A, genesT, Gs, S, W32p, shift);
    A.lab_base = base;
    A.n_perms = P; A.ppi = ppi; A.items_per_tile = (P + ppi - 1) / ppi; A.chunk_base = 0;
    A.unperm = unperm; A.hits = hits;
    const int T = sb::WALK_THREADS;
    const size_t words = (size_t)sb::walk_push_words(false, sb::WALK_NLAB) * (stack_units > 0 ? stack_units : 1) * T;
    std::vector<int> smem(words + GUARD, CANARY);
    const int64_t per_block = (int64_t)T * sb::WALK_NP;
    for (int64_t tile = 0; tile < (S + per_block - 1) / per_block; ++tile)
        for (int chunk = 0; chunk < A.items_per_tile; ++chunk)
            for (int tid = 0; tid < T; ++tid) {
                sb::blockIdx = {(int)tile, chunk, 0};
                sb::threadIdx = {tid, 0, 0};
                sb::sb_emul_shared = smem.data();
                sb::walk_permute_kernel<false>(A);
            }
    return guard_intact(smem, words) ? 0 : -2;
}

// K5, transposed launch: threads = labellings (labelsT [W32p][Ps], the label vectors transposed), constant rows =
// the walk-order bits of n_rows genes (rowsW [n_rows][W32p]); hits [n_rows][Ps] bytes, unperm [n_rows][3]
int emul_permute_transposed(const uint16_t *ops, int n_ops, const uint32_t *rowsW, int n_rows, int ppi,
                            const uint32_t *labelsT, int64_t Ps, int64_t P, int W32p, int shift, int stack_units,
                            const int32_t *unperm, uint8_t *hits)
{
    const int base = load_pool(ops, n_ops, rowsW, n_rows, W32p);
    if (base < 0 || ppi < 1 || ppi > sb::PERMS_PER_ITEM_MAX) return -1;
    sb::WalkArgs A;
    fill(A, labelsT, Ps, P, W32p, shift);
    A.S_total = Ps; A.lab_base = base; A.row_base = 0;
    A.n_perms = n_rows; A.ppi = ppi; A.items_per_tile = (n_rows + ppi - 1) / ppi;
    A.unperm = unperm; A.hits = hits;
    const int T = sb::WALK_THREADS;
    const size_t words = (size_t)sb::walk_push_words(false, 1) * (stack_units > 0 ? stack_units : 1) * T;
    std::vector<int> smem(words + GUARD, CANARY);
    const int64_t per_block = (int64_t)T * sb::WALK_NP;
    for (int64_t tile = 0; tile < (P + per_block - 1) / per_block; ++tile)
        for (int chunk = 0; chunk < A.items_per_tile; ++chunk)
            for (int tid = 0; tid < T; ++tid) {
                sb::blockIdx = {(int)tile, chunk, 0};
                sb::threadIdx = {tid, 0, 0};
                sb::sb_emul_shared = smem.data();
                sb::walk_permute_kernel<true>(A);
            }
    return guard_intact(smem, words) ? 0 : -2;
}

}  // extern "C"
