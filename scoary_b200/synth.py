"""Synthetic workloads of SURVEY.md 8(d) (BASELINE.json configs C2-C5).

presence M[g, j] ~ Bernoulli(f_g), f_g ~ U(0.02, 0.98); traits ~ Bernoulli(0.35);
10 planted causal genes per trait (gene = trait XOR Bernoulli(0.1)); optional
missing trait values; numpy default_rng(seed) (PCG64).  Trees: a seeded
random-join binary tree over the isolates (tree building is off the hot path).
"""
import numpy as np

from . import engine as eng
from . import tree as treemod

CONFIGS = {
    # name: (genes, isolates, traits, permutations, seed)
    "c2": (10_000, 1_000, 1, 0, 20260902),
    "c3": (50_000, 5_000, 1, 1_000, 20260903),
    "c4": (100_000, 10_000, 4, 10_000, 20260904),
    "c5": (1_000_000, 2_000, 1, 1_000, 20260905),
    "north_star": (50_000, 5_000, 1, 10_000, 20260903),
}


def isolate_names(n):
    return ["Iso_%05d" % i for i in range(n)]


def gene_names(g, start=0):
    return ["g%07d" % i for i in range(start, start + g)]


def make_traits(n_isolates, n_traits, seed, missing_frac=0.0):
    """int8 [T][N]: 1 / 0 / -1 (missing)."""
    rng = np.random.default_rng([seed, 0x7A17])
    t = (rng.random((n_traits, n_isolates)) < 0.35).astype(np.int8)
    if missing_frac > 0:
        miss = rng.random((n_traits, n_isolates)) < missing_frac
        t[miss] = -1
    return t


def _gene_chunk(lo, hi, n_isolates, seed, traits, planted, gene_offset):
    """uint8 [hi - lo][N]: rows lo .. hi - 1 of a shard; the random stream belongs to the chunk (seed, gene_offset + lo)."""
    rng = np.random.default_rng([seed, 0x6E6E, gene_offset + lo])
    f = rng.uniform(0.02, 0.98, size=(hi - lo, 1)).astype(np.float32)
    m = (rng.random((hi - lo, n_isolates), dtype=np.float32) < f).astype(np.uint8)
    if traits is not None and gene_offset == 0:
        for k in range(traits.shape[0]):
            for r in range(k * planted, (k + 1) * planted):
                if lo <= r < hi:
                    flip = (rng.random(n_isolates) < 0.1)
                    m[r - lo] = ((traits[k] == 1) ^ flip).astype(np.uint8)
    return m


def make_genes_packed(n_genes, n_isolates, seed, traits=None, planted=10, chunk=4096, gene_offset=0):
    """uint64 [G][W] bitset rows; rows [k*planted, (k+1)*planted) of shard 0 are planted for trait k."""
    W = eng.words_for(n_isolates)
    out = np.empty((n_genes, W), dtype=np.uint64)
    for lo in range(0, n_genes, chunk):
        hi = min(n_genes, lo + chunk)
        out[lo:hi] = eng.pack_rows(_gene_chunk(lo, hi, n_isolates, seed, traits, planted, gene_offset))
    return out


def make_genes_rows(lo, hi, n_genes, n_isolates, seed, traits=None, planted=10, chunk=4096):
    """Rows lo .. hi - 1 of make_genes_packed(n_genes, ...): one GPU's shard of a FIXED job (strong scaling), generated
    without the rest of the matrix (whole chunks are drawn, the shard is cut out of them)."""
    W = eng.words_for(n_isolates)
    out = np.empty((hi - lo, W), dtype=np.uint64)
    for c_lo in range(lo // chunk * chunk, hi, chunk):
        c_hi = min(n_genes, c_lo + chunk)
        rows = eng.pack_rows(_gene_chunk(c_lo, c_hi, n_isolates, seed, traits, planted, 0))
        a, b = max(lo, c_lo), min(hi, c_hi)
        out[a - lo:b - lo] = rows[a - c_lo:b - c_lo]
    return out


def make_tree(n_isolates, seed):
    rng = np.random.default_rng([seed, 0x7EE])
    return treemod.random_join_tree(isolate_names(n_isolates), rng)


def unpack_rows(bits_u64, n_isolates):
    """inverse of engine.pack_rows (for the oracle, small cases)."""
    b = np.ascontiguousarray(bits_u64, dtype=np.uint64)
    by = b.view(np.uint8).reshape(b.shape[0], -1)
    return np.unpackbits(by, axis=1, bitorder="little")[:, :n_isolates]
