"""TEST INFRASTRUCTURE: an oracle-backed stand-in for scoary_b200.engine.Engine, used ONLY by
the `-m "not gpu"` tests to exercise the host logic (collapse, Bonferroni/BH, sorting,
filtering, CSV text) against the reference's golden files without a GPU.  Fisher p comes
from SciPy here -- the arithmetic the reference itself uses -- so whole result files can be
compared as text.  Never importable from the product package."""
import numpy as np

from oracle import oracle as O
from scoary_b200 import synth


class FakeEngine:
    def __init__(self, device=0):
        self.G = self.N = self.W = 0
        self.traits, self.trees = {}, {}

    def set_genes(self, bits, n):
        self.m = synth.unpack_rows(bits, n)
        self.G, self.N, self.W = self.m.shape[0], n, bits.shape[1]

    def set_trait_vector(self, t, vec):
        self.traits[t] = np.asarray(vec, dtype=np.int8)

    def set_tree(self, t, left, right, leaf_to_col):
        self.trees[t] = (np.asarray(left), np.asarray(right), np.asarray(leaf_to_col))

    def upgma(self):
        return O.upgma_merges(self.m)

    def contingency_fisher(self, t, want_p=True, want_hash=False):
        counts = O.contingency(self.m, self.traits[t])
        p = O.fisher_scipy(counts) if want_p else None
        h = O.pattern_hash(self.m, self.traits[t]) if want_hash else None
        return counts, p, h

    def contingency_fisher_multi(self, t0, n_traits, want_p=True, want_hash=False):
        outs = [self.contingency_fisher(t0 + k, want_p, want_hash) for k in range(n_traits)]
        return (np.stack([o[0] for o in outs]), np.stack([o[1] for o in outs]) if want_p else None,
                np.stack([o[2] for o in outs]) if want_hash else None)

    def set_permute_mode(self, mode):
        pass

    def permute_range(self, t, perm_first, perm_count, seed=0, gene_idx=None):
        left, right, g, lab = self._walk_inputs(t, gene_idx)
        res = O.permute(left, right, g, lab, P=perm_first + perm_count, seed=seed & (2**64 - 1), trait=t, want_hits=True)
        return res["pairs"], res["hits"][:, perm_first:perm_first + perm_count].sum(axis=1).astype(np.int32)

    def _walk_inputs(self, t, gene_idx):
        left, right, cols = self.trees[t]
        rows = np.arange(self.G) if gene_idx is None else np.asarray(gene_idx)
        return left, right, self.m[np.ix_(rows, cols)], self.traits[t][cols].astype(np.uint8)

    def pairwise(self, t, gene_idx=None):
        left, right, g, lab = self._walk_inputs(t, gene_idx)
        return O.permute(left, right, g, lab, P=0)["pairs"]

    def permute(self, t, P, seed=0, gene_idx=None, early_stop=False, rmin=None):
        left, right, g, lab = self._walk_inputs(t, gene_idx)
        res = O.permute(left, right, g, lab, P=P, seed=seed & (2**64 - 1), trait=t, early_stop=early_stop)
        return res["pairs"], res["r"], res["n_done"]
