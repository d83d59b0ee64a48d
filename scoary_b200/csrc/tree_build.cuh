// tree_build.cuh -- SURVEY.md 8(f) rank 1: the UPGMA tree of the isolates, on the GPU.
//
// Replaces CreateTriangularDistanceMatrix (scoary/methods.py:619-644, scipy pdist 'hamming'),
// PopulateQuadTreeWithDistances (:646-665), upgma (:667-707) and classes.QuadTree (:97-196):
//   1. variable genes (present in some but not all allowed isolates, methods.py:496-497) are
//      transposed into an isolate-major bit matrix;
//   2. all-pairs Hamming distance = popcount(xor) / #variable genes (the diagonal is 1);
//   3. UPGMA: repeatedly merge the minimum cell.  The reference finds it by descending a
//      QuadTree, taking the smallest (value, i, j) in every quad from the coarsest level down:
//      that is the minimum cell with the smallest bit-interleaved (i, j) key, which is what the
//      argmin below uses.  Merged cluster keeps index i, sees retired clusters at distance 1,
//      its diagonal and everything of j become sys.maxsize (methods.py:686-703).
// Arithmetic is IEEE double with explicit round-to-nearest intrinsics (no FMA contraction), in
// the reference's operation order, so the merge order -- and therefore the tree -- is identical.
#pragma once
#include "common.cuh"

namespace sb {

constexpr double UPGMA_BIG = 9223372036854775807.0;   // float(sys.maxsize)

// ---- 1. variable-gene flags + transpose to isolate-major
// genes [G][W] uint64; allowed [W] (isolates taking part).  T [N][Gw] uint64: bit (g & 63) of
// word (g >> 6) of row j = gene g present in isolate j, for variable genes only.
// Block = 64 genes (one output word) x 256 isolates.
__global__ void __launch_bounds__(256) transpose_variable_kernel(const uint64_t *__restrict__ genes, int64_t G, int W,
                                                                 int N, const uint64_t *__restrict__ allowed,
                                                                 int n_allowed, int64_t Gw, uint64_t *__restrict__ T,
                                                                 unsigned long long *__restrict__ n_variable)
{
    __shared__ uint64_t s_rows[64][5];   // 64 genes x 4 words of isolates (+1 pad)
    __shared__ uint64_t s_var;           // bit g = gene variable
    __shared__ int s_cnt[64];
    const int64_t g0 = (int64_t)blockIdx.x * 64;
    const int w0 = blockIdx.y * 4;       // first isolate word of this block
    const int tid = threadIdx.x;
    // variable flag: every block recomputes it for its 64 genes over the full row (W is small)
    if (tid < 64) s_cnt[tid] = 0;
    if (tid == 0) s_var = 0;
    __syncthreads();
    for (int idx = tid; idx < 64 * W; idx += 256) {
        const int g = idx / W, w = idx - g * W;
        if (g0 + g < G) atomicAdd(&s_cnt[g], __popcll(genes[(g0 + g) * W + w] & allowed[w]));
    }
    __syncthreads();
    if (tid < 64) {
        const bool var = (g0 + tid < G) && s_cnt[tid] > 0 && s_cnt[tid] < n_allowed;
        if (var) atomicOr((unsigned long long *)&s_var, 1ULL << tid);
    }
    for (int idx = tid; idx < 64 * 4; idx += 256) {
        const int g = idx >> 2, w = idx & 3;
        s_rows[g][w] = (g0 + g < G && w0 + w < W) ? genes[(g0 + g) * W + w0 + w] : 0ULL;
    }
    __syncthreads();
    if (blockIdx.y == 0 && tid == 0) atomicAdd(n_variable, (unsigned long long)__popcll(s_var));
    const int j = w0 * 64 + tid;         // isolate of this thread
    if (j >= N) return;
    const uint64_t var = s_var;
    uint64_t word = 0;
    const int wl = tid >> 6, bl = tid & 63;
#pragma unroll 8
    for (int g = 0; g < 64; ++g) word |= ((s_rows[g][wl] >> bl) & 1ULL) << g;
    T[(int64_t)j * Gw + blockIdx.x] = word & var;
}

// ---- 2. all-pairs Hamming distances
// D [N][N] double, symmetric; D[i][i] = 1.  Block = 32 x 32 pairs, words streamed through
// shared memory in chunks of 32.
__global__ void __launch_bounds__(1024) hamming_kernel(const uint64_t *__restrict__ T, int N, int64_t Gw,
                                                       const unsigned long long *__restrict__ n_variable,
                                                       const uint64_t *__restrict__ allowed, double *__restrict__ D)
{
    __shared__ uint64_t sa[32][33], sb_[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;                 // upper triangle of tiles
    const int ti = threadIdx.y, tj = threadIdx.x;
    const int i = bi * 32 + ti, j = bj * 32 + tj;
    unsigned cnt = 0;
    for (int64_t c0 = 0; c0 < Gw; c0 += 32) {
        const int64_t c = c0 + tj;
        const int ri = bi * 32 + ti, rj = bj * 32 + ti;
        sa[ti][tj] = (ri < N && c < Gw) ? T[(int64_t)ri * Gw + c] : 0ULL;
        sb_[ti][tj] = (rj < N && c < Gw) ? T[(int64_t)rj * Gw + c] : 0ULL;
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) cnt += __popcll(sa[ti][k] ^ sb_[tj][k]);
        __syncthreads();
    }
    if (i < N && j < N) {
        const unsigned long long nv = *n_variable;
        const double den = (double)(nv > 0 ? nv : 1ULL);
        const bool ai = (allowed[i >> 6] >> (i & 63)) & 1ULL, aj = (allowed[j >> 6] >> (j & 63)) & 1ULL;
        double v = __ddiv_rn((double)cnt, den);
        if (i == j) v = 1.0;
        if (!ai || !aj) v = UPGMA_BIG;   // isolates that do not take part never merge
        D[(int64_t)i * N + j] = v;
        D[(int64_t)j * N + i] = v;
    }
}

// ---- 3. UPGMA
__device__ __forceinline__ unsigned long long morton_key(unsigned i, unsigned j)
{
    unsigned long long k = 0;
#pragma unroll
    for (int b = 15; b >= 0; --b) k = (k << 2) | (((unsigned long long)((i >> b) & 1u)) << 1) | ((j >> b) & 1u);
    return k;
}

struct UpgmaState {
    double *D;            // [N][N]
    double *rowmin_val;   // [N]
    int *rowmin_col;      // [N]
    double *size;         // [N]
    unsigned char *alive; // [N]
    unsigned char *redo;  // [N] rows whose minimum must be recomputed
    int *pick;            // [2] current (i, j); [2] = the merge step the next pick kernel performs
    int *merges;          // [N-1][2]
    int N;
};
// The three kernels of a merge step read the step counter pick[2] from device memory and do nothing once N - 1 merges
// are done, so a CUDA graph of UPGMA_GRAPH_STEPS steps can be replayed until the tree is complete (sb_upgma).
constexpr int UPGMA_GRAPH_STEPS = 64;

// minimum of row r by (value, column)
__device__ __forceinline__ void row_min_block(const UpgmaState &S, int r, double *s_val, int *s_col)
{
    const double *row = S.D + (int64_t)r * S.N;
    double bv = UPGMA_BIG * 4.0;
    int bc = 0x7fffffff;
    for (int c = threadIdx.x; c < S.N; c += blockDim.x) {
        const double v = row[c];
        if (v < bv || (v == bv && c < bc)) { bv = v; bc = c; }
    }
    s_val[threadIdx.x] = bv;
    s_col[threadIdx.x] = bc;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v = s_val[threadIdx.x + o];
            const int c = s_col[threadIdx.x + o];
            if (v < s_val[threadIdx.x] || (v == s_val[threadIdx.x] && c < s_col[threadIdx.x])) {
                s_val[threadIdx.x] = v;
                s_col[threadIdx.x] = c;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) upgma_rowmin_all_kernel(const UpgmaState S)
{
    __shared__ double s_val[256];
    __shared__ int s_col[256];
    const int r = blockIdx.x;
    row_min_block(S, r, s_val, s_col);
    if (threadIdx.x == 0) { S.rowmin_val[r] = s_val[0]; S.rowmin_col[r] = s_col[0]; S.redo[r] = 0; }
}

// global argmin over the row minima by (value, morton(i, j)); one block
__global__ void __launch_bounds__(1024) upgma_pick_kernel(const UpgmaState S)
{
    const int step = S.pick[2];
    if (step >= S.N - 1) return;
    __shared__ double s_val[1024];
    __shared__ unsigned long long s_key[1024];
    __shared__ int s_row[1024];
    double bv = UPGMA_BIG * 4.0;
    unsigned long long bk = ~0ULL;
    int br = -1;
    for (int r = threadIdx.x; r < S.N; r += 1024) {
        const double v = S.rowmin_val[r];
        const unsigned long long k = morton_key((unsigned)r, (unsigned)S.rowmin_col[r]);
        if (v < bv || (v == bv && k < bk)) { bv = v; bk = k; br = r; }
    }
    s_val[threadIdx.x] = bv; s_key[threadIdx.x] = bk; s_row[threadIdx.x] = br;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v = s_val[threadIdx.x + o];
            const unsigned long long k = s_key[threadIdx.x + o];
            if (v < s_val[threadIdx.x] || (v == s_val[threadIdx.x] && k < s_key[threadIdx.x])) {
                s_val[threadIdx.x] = v; s_key[threadIdx.x] = k; s_row[threadIdx.x] = s_row[threadIdx.x + o];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int i = s_row[0], j = S.rowmin_col[i];
        S.pick[0] = i; S.pick[1] = j;
        S.merges[2 * step] = i; S.merges[2 * step + 1] = j;
    }
}

// new row/column of cluster i, retire j, incremental row-minimum maintenance (grid over k)
__global__ void __launch_bounds__(256) upgma_update_kernel(const UpgmaState S)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= S.N || S.pick[2] >= S.N - 1) return;
    const int i = S.pick[0], j = S.pick[1], N = S.N;
    const double si = S.size[i], sj = S.size[j];
    const double ns = __dadd_rn(si, sj);
    double nd;
    if (k == i) nd = UPGMA_BIG;
    else if (!S.alive[k]) nd = 1.0;      // retired clusters: methods.py:687-688
    else nd = __ddiv_rn(__dadd_rn(__dmul_rn(S.D[(int64_t)i * N + k], si), __dmul_rn(S.D[(int64_t)j * N + k], sj)), ns);
    // (the reads above use the old rows i and j; the writes below touch column i/j of row k and
    //  entries k of rows i/j -- each thread reads D[i][k], D[j][k] before it overwrites them)
    const double vi = (k == j) ? UPGMA_BIG : nd;
    S.D[(int64_t)i * N + k] = vi;        // row i (entry j becomes BIG)
    S.D[(int64_t)k * N + i] = vi;        // column i
    S.D[(int64_t)j * N + k] = UPGMA_BIG; // row j
    S.D[(int64_t)k * N + j] = UPGMA_BIG; // column j
    // row-minimum maintenance for row k (rows i and j are recomputed in full)
    if (k == i || k == j) { S.redo[k] = 1; return; }
    const int mc = S.rowmin_col[k];
    if (mc == i || mc == j) { S.redo[k] = 1; return; }
    const double mv = S.rowmin_val[k];
    if (vi < mv || (vi == mv && i < mc)) { S.rowmin_val[k] = vi; S.rowmin_col[k] = i; }
}

__global__ void __launch_bounds__(256) upgma_redo_kernel(const UpgmaState S)
{
    __shared__ double s_val[256];
    __shared__ int s_col[256];
    const int r = blockIdx.x;
    if (S.pick[2] >= S.N - 1) return;
    if (!S.redo[r]) return;              // uniform per block
    row_min_block(S, r, s_val, s_col);
    if (threadIdx.x == 0) {
        S.rowmin_val[r] = s_val[0]; S.rowmin_col[r] = s_col[0]; S.redo[r] = 0;
    }
}

// End of a merge step (one thread): size / alive bookkeeping -- it must not race with upgma_update_kernel's reads, and
// nothing in upgma_redo_kernel reads it -- and the step counter, which every kernel of the step has read by now.
__global__ void upgma_finish_step_kernel(const UpgmaState S)
{
    if (S.pick[2] >= S.N - 1) return;
    const int i = S.pick[0], j = S.pick[1];
    S.size[i] = __dadd_rn(S.size[i], S.size[j]);
    S.size[j] = 0.0;
    S.alive[j] = 0;
    S.pick[2] += 1;
}

}  // namespace sb
