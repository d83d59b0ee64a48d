/*
 * scoary_b200.h -- C-ABI of libscoary_b200.so, the B200 (sm_100a) engine that
 * replaces the per-gene statistics / pairwise-comparison / permutation path of
 * Scoary.
 *
 * The reference (pure Python, /root/reference) has no FFI; the seam this
 * library sits behind is the set of call sites in scoary/methods.py:
 *   Setup_results(genedic, traitsdic, collapse)        methods.py:278 -> def :757
 *     Perform_statistics(traits, genes)                methods.py:798 -> def :930
 *     ss.fisher_exact(obs_table)                       methods.py:854
 *   PairWiseComparisons((domain, argdict))             methods.py:1094/1105 -> def :1208
 *     ConvertUPGMAtoPhyloTree(tree, GTC)               methods.py:1247 -> def :1386
 *       class PhyloTree / Tip                          scoary/classes.py:199-592
 *     Permute(tree, GTC, permutations, cutoffs)        methods.py:1305 -> def :1314
 *       PermuteGTC(GTC)                                methods.py:1350 -> def :1371
 * Each entry point below names the call site it replaces.  INTEGRATION.md shows
 * the ctypes binding a maintainer would add to the reference.
 *
 * Conventions: plain pointers and sizes only; the caller owns every buffer it
 * passes; the library never frees caller memory.  Every function returns 0 on
 * success and a negative sb_status on failure; sb_last_error() gives the text.
 * One context drives one GPU (one process per GPU; gene rows are sharded
 * across processes by the host).  A context is not re-entrant.  "host" entry
 * points block until their outputs are valid; "*_device" entry points only
 * enqueue work on the context's stream.
 *
 * There is no CPU fallback: without a CUDA device sb_create fails.
 */
#ifndef SCOARY_B200_H
#define SCOARY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_VERSION 100          /* 0.1.0 */
#define SB_MAX_TRAITS 64

typedef struct sb_ctx sb_ctx;

typedef enum {
    SB_OK = 0,
    SB_ERR_ARG = -1,            /* bad argument / call order */
    SB_ERR_CUDA = -2,           /* CUDA runtime error (text in sb_last_error) */
    SB_ERR_NOMEM = -3,
    SB_ERR_STATE = -4           /* genes / trait / tree not set */
} sb_status;

/* Counters since sb_create (or the last sb_stats_reset).  Kernel times are
 * only accumulated while profiling is on (sb_set_profiling), because they need
 * a pair of CUDA events around each launch. */
typedef struct {
    int64_t kernel_launches;        /* launches of THIS library's kernels */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    int64_t tests_contingency;      /* (gene, trait) tables built */
    int64_t tests_walks;            /* (gene, labelling) tree walks, incl. permutations */
    double ms_pack;                 /* K1 transpose/gather into walk order */
    double ms_fisher;               /* K2+K3 contingency + Fisher */
    double ms_shuffle;              /* Fisher-Yates label shuffles */
    double ms_walk;                 /* K4 unpermuted walks */
    double ms_permute;              /* K5 permutation walks (dominant kernel) */
    double ms_reduce;               /* hit-sequence reduction */
    int64_t launches_permute;       /* launches of the K5 kernel while profiling */
    int32_t sm_count;
    int32_t reserved;
    int64_t calls_transposed;       /* sb_permute calls that ran threads = labellings (few genes, many permutations) */
    double ms_epilogue;             /* device epilogue: sort, Bonferroni / Benjamini-Hochberg, binomial tests */
} sb_stats_t;

/* ---- life cycle -------------------------------------------------------- */
int sb_version(void);
/* device = CUDA ordinal.  Fails (SB_ERR_CUDA) when no usable GPU exists. */
int sb_create(int device, sb_ctx **out);
void sb_destroy(sb_ctx *ctx);
/* Text of the last error on ctx (ctx == NULL: last sb_create failure). */
const char *sb_last_error(const sb_ctx *ctx);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL
 * restores the context's own stream. */
int sb_set_stream(sb_ctx *ctx, void *cuda_stream);
int sb_synchronize(sb_ctx *ctx);
int sb_set_profiling(sb_ctx *ctx, int on);
/* K5 launch shape of sb_permute: 0 = chosen per call (default), 1 = threads are genes, 2 = threads are labellings
 * (the "few genes x many permutations" shape left after decideifbreak, scoary/methods.py:1022-1024, :1295-1310:
 * the DP is symmetric in the gene and the trait bit of a leaf, so the same walk runs with the roles exchanged).
 * Results are identical in every mode. */
int sb_set_permute_mode(sb_ctx *ctx, int mode);
int sb_stats(sb_ctx *ctx, sb_stats_t *out);
int sb_stats_reset(sb_ctx *ctx);

/* ---- inputs ------------------------------------------------------------ */
/* Gene presence/absence bitset, replaces genedic (methods.py:472-487; a cell
 * is "present" unless it is "", "0" or "-").  bits: row-major uint64[G][W],
 * bit (j & 63) of word (j >> 6) of row g = gene g present in isolate column j;
 * N isolates, W >= ceil(N/64) and W even (16-byte row pitch); bits at
 * positions >= N must be zero.  The host version copies to the device. */
int sb_set_genes(sb_ctx *ctx, const uint64_t *bits, int64_t G, int32_t N, int32_t W);
/* Same, but d_bits already lives in device memory and is borrowed, not copied;
 * it must stay valid until the next sb_set_genes* or sb_destroy. */
int sb_set_genes_device(sb_ctx *ctx, const uint64_t *d_bits, int64_t G, int32_t N, int32_t W);

/* Trait t (0 <= t < SB_MAX_TRAITS), replaces traitsdic[trait] (methods.py:546-614):
 * value bit = trait is "1", mask bit = isolate has a non-missing value for
 * this trait (missing isolates are never counted, methods.py:580-598).
 * Host pointers, uint64[W] each. */
int sb_set_trait(sb_ctx *ctx, int32_t t, const uint64_t *value, const uint64_t *mask);

/* Binary tree for trait t, already pruned of the trait's missing isolates
 * (PruneForMissing, methods.py:709-739).  n_internal nodes listed children
 * before parents, root last; child >= 0 is an internal node index, child < 0 is
 * leaf id ~child (0 .. n_internal); leaf_to_col[leaf id] = isolate column in
 * the gene bitset.  Replaces the nested-list `tree` argument of
 * ConvertUPGMAtoPhyloTree (methods.py:1386).  Host pointers.
 * Limit: at most 32 766 leaves (the DP keys hold pairs and supporting pairs in 28 bits).  Every binary tree up to
 * that size is accepted whatever its shape: the compiled walk program (<= 21 374 ops, a balanced tree) and at least
 * five label vectors fit the 62 KB constant pool; sb_set_tree itself rejects anything larger (SB_ERR_ARG). */
int sb_set_tree(sb_ctx *ctx, int32_t t, const int32_t *left, const int32_t *right,
                int32_t n_internal, const int32_t *leaf_to_col);

/* ---- the hot path ------------------------------------------------------ */
/* Setup_results inner loop (methods.py:791-857): for every gene, the 2x2 table
 * of Perform_statistics (methods.py:930-982) and the two-sided Fisher exact p
 * of ss.fisher_exact (methods.py:854, SciPy 1.18.1 rule).
 *   counts int32[G][4] = tpgp, tngp, tpgn, tngn          (bit-exact)
 *   p      double[G]   two-sided p (<= 1e-10 relative vs SciPy); 1.0 when a
 *                      margin is zero (the host applies the skip rule of
 *                      methods.py:804-814 from the counts)
 *   hash   uint64[G][2] 128-bit hash of the row masked by the trait's mask,
 *                      for --collapse grouping (methods.py:823-840); may be NULL
 * Any output pointer may be NULL. */
int sb_contingency_fisher(sb_ctx *ctx, int32_t t, int32_t *counts, double *p, uint64_t *hash);
int sb_contingency_fisher_device(sb_ctx *ctx, int32_t t, int32_t *d_counts, double *d_p,
                                 uint64_t *d_hash);
/* The same for the traits in slots t0 .. t0 + n_traits - 1 in ONE pass over the gene rows: the
 * reference loops traits outside genes (methods.py:771, :791) and so reads every gene once per
 * trait; here the trait vectors are staged in shared memory and each row is read once (C4: 340
 * instead of 1 288 bytes per test).  Outputs are trait-major: counts int32[n_traits][G][4],
 * p double[n_traits][G], hash uint64[n_traits][G][2]; same values as n_traits single calls. */
int sb_contingency_fisher_multi(sb_ctx *ctx, int32_t t0, int32_t n_traits, int32_t *counts, double *p,
                                uint64_t *hash);
int sb_contingency_fisher_multi_device(sb_ctx *ctx, int32_t t0, int32_t n_traits, int32_t *d_counts,
                                       double *d_p, uint64_t *d_hash);

/* ConvertUPGMAtoPhyloTree (methods.py:1386-1402) for S genes: max contrasting
 * pairs and, given those, max supporting / opposing pairs (classes.py:199-592).
 * gene_idx: int64[S] gene rows (NULL = rows 0..S-1).  pairs int32[S][3] =
 * Total, Pro, Anti -- bit-exact. */
int sb_pairwise(sb_ctx *ctx, int32_t t, const int64_t *gene_idx, int64_t S, int32_t *pairs);
int sb_pairwise_device(sb_ctx *ctx, int32_t t, const int64_t *d_gene_idx, int64_t S,
                       int32_t *d_pairs);

/* Permute (methods.py:1314-1369) for S genes and P label permutations.
 * Permutation i of trait t is a Fisher-Yates shuffle of the trait's labels
 * (PermuteGTC, methods.py:1371-1384) driven by Philox4x32-10 keyed by seed with
 * counter (step/2, i, t, 0x5C0A27) -- see DESIGN.md; the reference's own stream
 * is an unseeded Mersenne Twister, so agreement with it is distributional.
 * hit_i = (S_i/Total_i >= S/Total) with S = Pro if Pro >= Anti else Anti
 * (methods.py:1333-1355), evaluated exactly in integers.
 *   early_stop = 0: all P permutations count; r = sum of hits, n_done = P.
 *   early_stop = 1: the reference's sequential rule (methods.py:1360-1363) on
 *     the ordered hit sequence: first i >= 30 with r_i >= rmin[i] stops,
 *     n_done = i + 1.  rmin int32[P] is built by the host from the same
 *     1 - binom.cdf(r, i, 0.1) < 0.05 test.
 * Empirical p = (r + 1) / (n_done + 1) is formed by the host.
 *   pairs int32[S][3] unpermuted Total, Pro, Anti (may be NULL);
 *   r int32[S]; n_done int32[S].
 * With early_stop = 1 the permutations are walked in rounds and the library waits for each
 * round's survivor count, so even the *_device variant synchronises the stream internally. */
int sb_permute(sb_ctx *ctx, int32_t t, const int64_t *gene_idx, int64_t S, int32_t P,
               uint64_t seed, int32_t early_stop, const int32_t *rmin, int32_t *pairs,
               int32_t *r, int32_t *n_done);
int sb_permute_device(sb_ctx *ctx, int32_t t, const int64_t *d_gene_idx, int64_t S, int32_t P,
                      uint64_t seed, int32_t early_stop, const int32_t *d_rmin,
                      int32_t *d_pairs, int32_t *d_r, int32_t *d_n_done);

/* Test hook: the permuted label vectors themselves, by leaf id.
 * labels uint8[P][n_leaves] (host). */
int sb_debug_shuffled_labels(sb_ctx *ctx, int32_t t, int32_t P, uint64_t seed, uint8_t *labels);

/* ---- tree construction ------------------------- */
/* UPGMA tree of the N isolates from the gene bitset set with sb_set_genes*: relative Hamming
 * distances over the variable genes (CreateTriangularDistanceMatrix, methods.py:619-644),
 * then upgma (methods.py:667-707) with the QuadTree's argmin tie-break (classes.py:155-196).
 * merges int32[N-1][2]: step s joins clusters (i, j); the joined cluster keeps index i
 * (methods.py:700-703), so the host rebuilds the nested list with
 * cluster[i] = [cluster[i], cluster[j]].  The merge order is identical to the reference's. */
int sb_upgma(sb_ctx *ctx, int32_t *merges);

/* Exhaustive Permute for the permutations perm_first .. perm_first + perm_count - 1 of a job: r[s] = hits among those
 * labellings only (the labelling of permutation i depends on (seed, trait slot, i) alone, so the hit counts of
 * disjoint ranges add up to sb_permute's r for the whole job).  This is how N GPUs split a job whose gene shards would
 * be too small to fill them: every GPU walks all S genes under its own range of the permutations and the r vectors
 * are summed (one all-reduce).  pairs (may be NULL) = the unpermuted Total, Pro, Anti as in sb_permute. */
int sb_permute_range(sb_ctx *ctx, int32_t t, const int64_t *gene_idx, int64_t S, int32_t perm_first,
                     int32_t perm_count, uint64_t seed, int32_t *pairs, int32_t *r);
int sb_permute_range_device(sb_ctx *ctx, int32_t t, const int64_t *d_gene_idx, int64_t S, int32_t perm_first,
                            int32_t perm_count, uint64_t seed, int32_t *d_pairs, int32_t *d_r);

/* ---- epilogue on the device ---- */
/* What Setup_results does with the p-values of one trait after the gene loop (methods.py:900-925) and the sort
 * every later step relies on (SortResultsAndSetKey, methods.py:1448-1454), for the 1M-row scale:
 *   keep   uint8[n]   1 = the gene was tested (skip rule methods.py:804-814 applied by the caller); the _device
 *                     variant derives it from counts int32[n][4] (tpgp+tngp > 0 and tpgn+tngn > 0)
 *   n_tests           the reference's number_of_tests; <= 0: the number of tested genes
 *   order  int32[n]   first *n_tested entries: tested genes by ascending p, ties in gene order (stable, as sorted())
 *   bonferroni, bh    double[n] by gene: min(p m, 1) and the Benjamini-Hochberg step-up value with the reference's
 *                     tie rule (a gene tied with its less significant neighbour inherits that neighbour's value,
 *                     methods.py:908-919), min(., 1); NaN for untested genes.  Same operations in the same order as
 *                     the reference (one multiplication, one division): bit-identical to it.
 * Any output may be NULL. */
int sb_adjust_pvalues(sb_ctx *ctx, const double *p, const uint8_t *keep, int64_t n, int64_t n_tests,
                      int32_t *order, double *bonferroni, double *bh, int64_t *n_tested);
int sb_adjust_pvalues_device(sb_ctx *ctx, const double *d_p, const int32_t *d_counts, int64_t n,
                             int64_t n_tests, int32_t *d_order, double *d_bonferroni, double *d_bh,
                             int64_t *n_tested);
/* ss.binom_test(k, n, 0.5) of PairWiseComparisons (methods.py:1267-1275; SciPy binomtest, two-sided) for `count`
 * (k, n) pairs: p = 1 if 2k == n, else min(1, 2 P[X <= min(k, n-k)]); NaN where n == 0.  <= 1e-13 relative. */
int sb_binom_two_sided(sb_ctx *ctx, const int32_t *k, const int32_t *n, int64_t count, double *p);

/* ---- input packing (host code, no GPU needed) ---- */
/* The per-cell loop of Csv_to_dic_Roary (methods.py:445-497) done natively on the raw bytes of
 * the gene presence/absence file.  Dialect: csv.reader(skipinitialspace=True, delimiter=d) as
 * the reference opens it (methods.py:350-351).
 * sb_csv_row_starts: byte offsets of the data rows (all rows after the header row); pass
 *   row_starts = NULL to count.  Returns the number of data rows or -1.  A '"' opens a quoted
 *   field only at the start of a field (csv.reader's rule): `5" nuclease` is plain text.
 * sb_csv_pack_rows: for each data row, bit i of the output row = cell of column keep_cols[i] is
 *   present (not "", "0" or "-", methods.py:476-487); bits uint64[n_rows][W].  lead_ranges
 *   [n_rows][n_lead][2] receives the byte range of the fields lead_cols[] (identifier,
 *   annotation, ...); begin < 0 marks a field with escaped quotes that the host must unescape
 *   (its range starts at -begin - 1).  row_fields[n_rows] = number of fields in the row.
 *   Returns 0, or -(row + 1) of the first row too short for the requested columns.
 * sb_csv_gather_fields: the text of lead field k of every row copied back to back into out (offsets[n_rows + 1];
 *   fields with begin < 0 contribute nothing); out = NULL only fills the offsets.  Returns the byte count or -1. */
int64_t sb_csv_row_starts(const char *buf, int64_t len, char delimiter, int64_t *row_starts,
                          int64_t max_rows, int64_t *header_end);
/* sb_csv_scan_rows: the same in one pass -- *row_starts_out receives a malloc'ed array (NULL without rows) that
 * sb_csv_free releases; returns the number of data rows or -1. */
int64_t sb_csv_scan_rows(const char *buf, int64_t len, char delimiter, int64_t **row_starts_out,
                         int64_t *header_end);
void sb_csv_free(void *p);
int64_t sb_csv_pack_rows(const char *buf, int64_t len, char delimiter, const int64_t *row_starts,
                         int64_t n_rows, const int32_t *keep_cols, int32_t n_keep, uint64_t *bits,
                         int32_t W, const int32_t *lead_cols, int32_t n_lead, int64_t *lead_ranges,
                         int32_t *row_fields);
int64_t sb_csv_gather_fields(const char *buf, const int64_t *lead_ranges, int64_t n_rows, int32_t n_lead,
                             int32_t k, char *out, int64_t out_cap, int64_t *offsets);

/* ---- VCF input (host code, no GPU needed) ---- */
/* scoary/vcf2scoary.py:50-218 followed by the cell loop of Csv_to_dic_Roary (methods.py:445-497),
 * done natively on the raw bytes of a VCF 4.x file: one output row per ALT allele, genotype = the
 * sample cell up to its first ':', single-ALT lines present unless "", "0" or "-", multi-ALT
 * lines present where int(genotype) is the allele's number ("." = absent), optional TYPE= filter.
 * sb_vcf_line_starts: byte offsets of the non-empty lines from the "#CHROM" header on (the "##"
 *   lines before it are skipped); line_starts = NULL to count.  *needs_python = 1 if the header
 *   or a variant line holds a '"' or a byte >= 0x80 (csv quoting / Unicode rules could matter:
 *   use the Python parser).
 * sb_vcf_count_rows: rows_per_line[i] for the variant lines (not the header); `types` is a
 *   '\n'-separated list of wanted TYPE values or NULL for all.  Returns the total number of rows,
 *   or -(line + 1) for a line with fewer than nine fields.
 * sb_vcf_pack_rows: bits uint64[n_rows][W]; keep[n_samples] = bit index of each sample column or
 *   -1 to drop it; ranges[n_rows][3][2] = byte ranges of CHROM, POS, ID; line_of[n_rows] = source
 *   line.  Returns 0, or -(line + 1) for the first line that must go to the Python parser (sample
 *   count differs from the header, or a multi-ALT genotype that is neither "." nor plain digits). */
int64_t sb_vcf_line_starts(const char *buf, int64_t len, int64_t *line_starts, int64_t max_lines,
                           int32_t *needs_python);
int64_t sb_vcf_count_rows(const char *buf, int64_t len, const int64_t *line_starts, int64_t n_lines,
                          const char *types, int64_t types_len, int32_t *rows_per_line);
int64_t sb_vcf_pack_rows(const char *buf, int64_t len, const int64_t *line_starts, int64_t n_lines,
                         const int32_t *rows_per_line, const int64_t *row_offset, const char *types,
                         int64_t types_len, const int32_t *keep, int32_t n_samples, uint64_t *bits,
                         int32_t W, int64_t *ranges, int64_t *line_of);

/* Test hook, host only (no GPU needed): the tree compiler behind sb_set_tree.  Writes the stack
 * program (uint16 ops: low 4 bits = op type, high 12 bits = count; see csrc/walk.cuh) and the
 * order in which the program consumes the leaves.  Returns the number of ops or a negative
 * sb_status (text in sb_last_error(NULL)). */
int sb_debug_compile_tree(const int32_t *left, const int32_t *right, int32_t n_internal,
                          uint16_t *ops_out, int32_t max_ops, int32_t *leaf_of_pos,
                          int32_t *stack_units);

/* The same hook for builds that pad the leaf stream (-DSB_WALK_PADDED=1, an experimental kernel variant):
 * leaf_of_pos [max_pos] receives the stream, -1 marking a pad position, and *n_pos its length
 * (= n_internal + 1 in the default build). */
int sb_debug_compile_tree2(const int32_t *left, const int32_t *right, int32_t n_internal,
                           uint16_t *ops_out, int32_t max_ops, int32_t *leaf_of_pos, int32_t max_pos,
                           int32_t *n_pos, int32_t *stack_units);

/* Integer-pipe microbenchmark used for the walk kernels' roofline denominator:
 * runs `iters` rounds of dependent add/max chains on every SM and returns the
 * measured int32 add+max operations per second. */
int sb_int32_peak(sb_ctx *ctx, int32_t iters, double *ops_per_s);
/* Design probe (tools/pipes.py): warp instructions per second of eight instruction kinds the walk kernels are made of
 * (VIADDMNMX, VIADDMNMX.S16x2, VIMNMX3.S16x2, VIADD.16x2, LOP3+SHF, VIMNMX3, ISETP+SEL, IMAD); out8 double[8]. */
int sb_debug_pipe_rates(sb_ctx *ctx, int32_t iters, double *out8);

#ifdef __cplusplus
}
#endif
#endif /* SCOARY_B200_H */
