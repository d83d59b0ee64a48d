"""Array-level host API over the C-ABI (include/scoary_b200.h).

Everything that touches numbers on the hot path happens inside
libscoary_b200.so on the GPU; this module only packs bits, flattens trees and
moves numpy buffers across ctypes.
"""
import ctypes

import numpy as np

from . import _lib
from . import tree as treemod


def words_for(n_isolates):
    """uint64 words per row: ceil(N/64) rounded up to even (16-byte pitch)."""
    w = (int(n_isolates) + 63) // 64
    return w + (w & 1)


def pack_rows(matrix_u8):
    """uint8/bool [G][N] presence matrix -> uint64 [G][W] bitset rows
    (bit j&63 of word j>>6 = column j)."""
    m = np.ascontiguousarray(matrix_u8, dtype=np.uint8)
    if m.ndim == 1:
        m = m[None, :]
    G, N = m.shape
    W = words_for(N)
    pad = W * 64 - N
    if pad:
        m = np.concatenate([m, np.zeros((G, pad), dtype=np.uint8)], axis=1)
    packed = np.packbits(m, axis=1, bitorder="little")          # [G][W*8] bytes
    return np.ascontiguousarray(packed).view(np.uint64).reshape(G, W)


def pack_trait(vec_i8):
    """int8 [N] with 1 / 0 / -1 (missing) -> (value bits, mask bits), uint64 [W]."""
    v = np.asarray(vec_i8, dtype=np.int8)
    value = pack_rows((v == 1).astype(np.uint8))[0]
    mask = pack_rows((v >= 0).astype(np.uint8))[0]
    return value, mask


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class EngineError(RuntimeError):
    pass


class Engine:
    """One context = one GPU (one process per GPU)."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        self._ctx = ctypes.c_void_p()
        rc = self._lib.sb_create(int(device), ctypes.byref(self._ctx))
        if rc != 0:
            msg = self._lib.sb_last_error(None)
            raise EngineError("sb_create failed (%d): %s" % (rc, msg.decode() if msg else "?"))
        self.device = int(device)
        self.G = self.N = self.W = 0
        self._keep = {}

    # -- plumbing
    def _check(self, rc):
        if rc != 0:
            msg = self._lib.sb_last_error(self._ctx)
            raise EngineError("libscoary_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.sb_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_handle):
        self._check(self._lib.sb_set_stream(self._ctx, ctypes.c_void_p(int(cuda_stream_handle or 0))))

    def synchronize(self):
        self._check(self._lib.sb_synchronize(self._ctx))

    def set_profiling(self, on):
        self._check(self._lib.sb_set_profiling(self._ctx, 1 if on else 0))

    def stats(self):
        st = _lib.SbStats()
        self._check(self._lib.sb_stats(self._ctx, ctypes.byref(st)))
        return {name: getattr(st, name) for name, _ in _lib.SbStats._fields_ if name != "reserved"}

    def stats_reset(self):
        self._check(self._lib.sb_stats_reset(self._ctx))

    def int32_peak(self, iters=4096):
        out = ctypes.c_double()
        self._check(self._lib.sb_int32_peak(self._ctx, int(iters), ctypes.byref(out)))
        return out.value

    # -- inputs
    def set_genes(self, bits_u64, n_isolates):
        b = np.ascontiguousarray(bits_u64, dtype=np.uint64)
        G, W = b.shape
        self._check(self._lib.sb_set_genes(self._ctx, _ptr(b), G, int(n_isolates), W))
        self._keep["genes"] = b
        self.G, self.N, self.W = G, int(n_isolates), W

    def set_genes_device(self, dev_ptr, G, n_isolates, W):
        self._check(self._lib.sb_set_genes_device(self._ctx, ctypes.c_void_p(int(dev_ptr)), int(G), int(n_isolates),
                                                  int(W)))
        self.G, self.N, self.W = int(G), int(n_isolates), int(W)

    def set_genes_matrix(self, matrix_u8):
        m = np.asarray(matrix_u8)
        self.set_genes(pack_rows(m), m.shape[1])

    def set_trait(self, t, value_bits, mask_bits):
        v = np.ascontiguousarray(value_bits, dtype=np.uint64)
        m = np.ascontiguousarray(mask_bits, dtype=np.uint64)
        if v.shape != (self.W,) or m.shape != (self.W,):
            raise ValueError("trait bit vectors must have %d words" % self.W)
        self._check(self._lib.sb_set_trait(self._ctx, int(t), _ptr(v), _ptr(m)))

    def set_trait_vector(self, t, vec_i8):
        v = np.asarray(vec_i8, dtype=np.int8)
        if v.shape != (self.N,):
            raise ValueError("trait vector must have %d entries" % self.N)
        value, mask = pack_trait(v)
        self.set_trait(t, value, mask)

    def set_tree(self, t, left, right, leaf_to_col):
        l = np.ascontiguousarray(left, dtype=np.int32)
        r = np.ascontiguousarray(right, dtype=np.int32)
        c = np.ascontiguousarray(leaf_to_col, dtype=np.int32)
        if len(l) != len(r) or len(c) != len(l) + 1:
            raise ValueError("tree arrays have inconsistent lengths")
        self._check(self._lib.sb_set_tree(self._ctx, int(t), _ptr(l), _ptr(r), len(l), _ptr(c)))

    def set_tree_nested(self, t, nested, column_of):
        """nested-list tree + {isolate name: gene-bitset column}."""
        left, right, names = treemod.flatten(nested)
        cols = np.asarray([column_of[n] for n in names], dtype=np.int32)
        self.set_tree(t, left, right, cols)
        return names

    # -- tree construction
    def upgma(self):
        """UPGMA merge list int32 [N-1][2] over the isolates of the current gene bitset."""
        merges = np.empty((self.N - 1, 2), dtype=np.int32)
        self._check(self._lib.sb_upgma(self._ctx, _ptr(merges)))
        return merges

    # -- epilogue
    def adjust_pvalues(self, p, keep, n_tests=0):
        """Setup_results' Bonferroni / Benjamini-Hochberg columns and the p-sort (methods.py:900-925, :1448-1454):
        -> (order int32 [m] tested genes by ascending p, bonferroni f64 [n], bh f64 [n])."""
        p = np.ascontiguousarray(p, dtype=np.float64)
        keep = np.ascontiguousarray(keep, dtype=np.uint8)
        n = len(p)
        order = np.empty(n, dtype=np.int32)
        bonf = np.empty(n, dtype=np.float64)
        bh = np.empty(n, dtype=np.float64)
        m = ctypes.c_int64()
        self._check(self._lib.sb_adjust_pvalues(self._ctx, _ptr(p), _ptr(keep), n, int(n_tests), _ptr(order), _ptr(bonf),
                                                _ptr(bh), ctypes.byref(m)))
        return order[:m.value], bonf, bh

    def adjust_pvalues_device(self, p_ptr, counts_ptr, n, order_ptr, bonf_ptr, bh_ptr, n_tests=0):
        m = ctypes.c_int64()
        self._check(self._lib.sb_adjust_pvalues_device(self._ctx, ctypes.c_void_p(int(p_ptr)), ctypes.c_void_p(int(counts_ptr)),
                                                       int(n), int(n_tests), ctypes.c_void_p(int(order_ptr) or None),
                                                       ctypes.c_void_p(int(bonf_ptr) or None),
                                                       ctypes.c_void_p(int(bh_ptr) or None), ctypes.byref(m)))
        return m.value

    def binom_two_sided(self, k, n):
        """ss.binom_test(k, n, 0.5) for arrays (methods.py:1267-1275)."""
        k = np.ascontiguousarray(k, dtype=np.int32).reshape(-1)
        n = np.ascontiguousarray(n, dtype=np.int32).reshape(-1)
        out = np.empty(len(k), dtype=np.float64)
        self._check(self._lib.sb_binom_two_sided(self._ctx, _ptr(k), _ptr(n), len(k), _ptr(out)))
        return out

    # -- hot path, host buffers
    def contingency_fisher(self, t, want_p=True, want_hash=False):
        G = self.G
        counts = np.empty((G, 4), dtype=np.int32)
        p = np.empty(G, dtype=np.float64) if want_p else None
        h = np.empty((G, 2), dtype=np.uint64) if want_hash else None
        self._check(self._lib.sb_contingency_fisher(self._ctx, int(t), _ptr(counts), _ptr(p), _ptr(h)))
        return counts, p, h

    def contingency_fisher_multi(self, t0, n_traits, want_p=True, want_hash=False):
        """Traits t0 .. t0 + n_traits - 1 in one pass over the gene rows (every row is read once for all of them):
        counts [T][G][4], p [T][G], hash [T][G][2]."""
        G, T = self.G, int(n_traits)
        counts = np.empty((T, G, 4), dtype=np.int32)
        p = np.empty((T, G), dtype=np.float64) if want_p else None
        h = np.empty((T, G, 2), dtype=np.uint64) if want_hash else None
        self._check(self._lib.sb_contingency_fisher_multi(self._ctx, int(t0), T, _ptr(counts), _ptr(p), _ptr(h)))
        return counts, p, h

    def set_permute_mode(self, mode):
        """K5 launch shape: 0 chosen per call, 1 threads = genes, 2 threads = labellings (same results)."""
        self._check(self._lib.sb_set_permute_mode(self._ctx, int(mode)))

    def pairwise(self, t, gene_idx=None):
        idx = None if gene_idx is None else np.ascontiguousarray(gene_idx, dtype=np.int64)
        S = self.G if idx is None else len(idx)
        pairs = np.empty((S, 3), dtype=np.int32)
        if S == 0:
            return pairs
        self._check(self._lib.sb_pairwise(self._ctx, int(t), _ptr(idx), S, _ptr(pairs)))
        return pairs

    def permute(self, t, P, seed=0, gene_idx=None, early_stop=False, rmin=None):
        idx = None if gene_idx is None else np.ascontiguousarray(gene_idx, dtype=np.int64)
        S = self.G if idx is None else len(idx)
        pairs = np.empty((S, 3), dtype=np.int32)
        r = np.empty(S, dtype=np.int32)
        nd = np.empty(S, dtype=np.int32)
        if S == 0:
            return pairs, r, nd
        rm = None
        if early_stop:
            if rmin is None:
                raise ValueError("early_stop needs the rmin table")
            rm = np.ascontiguousarray(rmin, dtype=np.int32)
            if len(rm) < P:
                raise ValueError("rmin must have P entries")
        self._check(self._lib.sb_permute(self._ctx, int(t), _ptr(idx), S, int(P), int(seed) & (2**64 - 1),
                                         1 if early_stop else 0, _ptr(rm), _ptr(pairs), _ptr(r), _ptr(nd)))
        return pairs, r, nd

    def permute_range(self, t, perm_first, perm_count, seed=0, gene_idx=None):
        """Exhaustive Permute over permutations perm_first .. perm_first + perm_count - 1 of a job:
        (pairs, r) with r = hits among those labellings (the ranges of a job add up to permute()'s r)."""
        idx = None if gene_idx is None else np.ascontiguousarray(gene_idx, dtype=np.int64)
        S = self.G if idx is None else len(idx)
        pairs = np.empty((S, 3), dtype=np.int32)
        r = np.empty(S, dtype=np.int32)
        if S == 0:
            return pairs, r
        self._check(self._lib.sb_permute_range(self._ctx, int(t), _ptr(idx), S, int(perm_first), int(perm_count),
                                               int(seed) & (2**64 - 1), _ptr(pairs), _ptr(r)))
        return pairs, r

    def permute_range_device(self, t, S, perm_first, perm_count, seed, pairs_ptr, r_ptr, gene_idx_ptr=0):
        self._check(self._lib.sb_permute_range_device(self._ctx, int(t), ctypes.c_void_p(int(gene_idx_ptr) or None), int(S),
                                                      int(perm_first), int(perm_count), int(seed) & (2**64 - 1),
                                                      ctypes.c_void_p(int(pairs_ptr) or None), ctypes.c_void_p(int(r_ptr))))

    def shuffled_labels(self, t, P, seed, n_leaves):
        out = np.empty((P, n_leaves), dtype=np.uint8)
        self._check(self._lib.sb_debug_shuffled_labels(self._ctx, int(t), int(P), int(seed) & (2**64 - 1), _ptr(out)))
        return out

    # -- hot path, device buffers (pointers as ints, e.g. torch tensor.data_ptr()); enqueue only
    def contingency_fisher_device(self, t, counts_ptr, p_ptr, hash_ptr=0):
        self._check(self._lib.sb_contingency_fisher_device(self._ctx, int(t), ctypes.c_void_p(int(counts_ptr) or None),
                                                           ctypes.c_void_p(int(p_ptr) or None),
                                                           ctypes.c_void_p(int(hash_ptr) or None)))

    def contingency_fisher_multi_device(self, t0, n_traits, counts_ptr, p_ptr, hash_ptr=0):
        self._check(self._lib.sb_contingency_fisher_multi_device(self._ctx, int(t0), int(n_traits),
                                                                 ctypes.c_void_p(int(counts_ptr) or None),
                                                                 ctypes.c_void_p(int(p_ptr) or None),
                                                                 ctypes.c_void_p(int(hash_ptr) or None)))

    def pairwise_device(self, t, S, pairs_ptr, gene_idx_ptr=0):
        self._check(self._lib.sb_pairwise_device(self._ctx, int(t), ctypes.c_void_p(int(gene_idx_ptr) or None), int(S),
                                                 ctypes.c_void_p(int(pairs_ptr))))

    def permute_device(self, t, S, P, seed, pairs_ptr, r_ptr, n_done_ptr, gene_idx_ptr=0, early_stop=False, rmin_ptr=0):
        self._check(self._lib.sb_permute_device(self._ctx, int(t), ctypes.c_void_p(int(gene_idx_ptr) or None), int(S),
                                                int(P), int(seed) & (2**64 - 1), 1 if early_stop else 0,
                                                ctypes.c_void_p(int(rmin_ptr) or None),
                                                ctypes.c_void_p(int(pairs_ptr) or None), ctypes.c_void_p(int(r_ptr)),
                                                ctypes.c_void_p(int(n_done_ptr))))


def shard_bounds(n_items, world_size):
    """Contiguous gene-row blocks, one per rank."""
    base, rem = divmod(int(n_items), int(world_size))
    bounds, start = [], 0
    for r in range(world_size):
        size = base + (1 if r < rem else 0)
        bounds.append((start, start + size))
        start += size
    return bounds
