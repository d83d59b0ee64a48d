"""TEST INFRASTRUCTURE: size-independent properties of the hot path, checked at sizes where the
oracle cannot redo the whole job (BASELINE.json's C3: 50 000 genes x 5 000 isolates x 1 000
permutations) -- plus a direct oracle comparison on a handful of genes at the FULL isolate and
permutation counts.  The same checker runs on the CPU against the oracle-backed FakeEngine at a
small size (tests/test_full_size_properties_cpu.py), which is what validates the checker itself.

Properties (all follow from the reference's definitions, none from this implementation):
  counts      the four cells are the popcounts of gene & trait & mask etc. (methods.py:930-982)
  Fisher      two-sided p is unchanged when the gene row is complemented (the table's rows swap)
  walk        complementing the gene turns every supporting pair (AB+ab) into an opposing pair
              (aB+Ab) and back: Total stays, Pro and Anti swap (classes.py:459-572); Total cannot
              exceed the rarer gene state or the rarer trait state among the leaves
  permutation exhaustive mode runs all P labellings; labellings depend on (seed, trait, index)
              only, so any subset of genes gives the same hit counts as the full run, and the
              complemented gene gives the same hit count whenever the unpermuted Pro != Anti
              (the tested side swaps with the statistic, methods.py:1333-1355)
"""
import numpy as np

from oracle import oracle as O
from scoary_b200 import engine as eng
from scoary_b200 import synth
from scoary_b200 import tree as treemod

FISHER_RTOL = 1e-10


def _popcount_rows(words):
    return np.bitwise_count(words).sum(axis=1).astype(np.int64)


def check(engine, G, N, P, seed, missing=0.0, n_oracle=8, n_subset=1000, T=1):
    """T > 1: the job has T traits (BASELINE C4 has four).  The Fisher pass of all of them runs as ONE multi-trait call
    and must equal the single-trait calls bit for bit; the walk / permutation properties are then checked for every
    trait in turn (its own mask, pruned tree and labellings)."""
    traits = synth.make_traits(N, T, seed, missing_frac=missing)
    bits = synth.make_genes_packed(G, N, seed, traits=traits)
    out = None
    if T > 1:
        engine.set_genes(bits, N)
        for t in range(T):
            engine.set_trait_vector(t, traits[t])
        c_all, p_all, h_all = engine.contingency_fisher_multi(0, T, want_hash=True)
        for t in range(T):
            c1, p1, h1 = engine.contingency_fisher(t, want_hash=True)
            assert np.array_equal(c_all[t], c1) and np.array_equal(h_all[t], h1)
            assert np.array_equal(p_all[t].view(np.uint64), p1.view(np.uint64))
    for t in range(T):
        res = _check_trait(engine, bits, traits[t], t, G, N, P, seed, n_oracle, n_subset)
        out = res if out is None else {k: out[k] + res[k] for k in out}
    return out


def _check_trait(engine, bits, vec, trait_index, G, N, P, seed, n_oracle, n_subset):
    value, mask = eng.pack_trait(vec)
    names = synth.isolate_names(N)
    nested = treemod.prune(synth.make_tree(N, seed), [names[j] for j in range(N) if vec[j] < 0])
    left, right, order = treemod.flatten(nested)
    col = {n: j for j, n in enumerate(names)}
    cols = np.asarray([col[n] for n in order], dtype=np.int32)
    n_mask, n_pos = int(np.bitwise_count(mask).sum()), int(np.bitwise_count(value & mask).sum())

    engine.set_genes(bits, N)
    engine.set_trait_vector(trait_index, vec)
    engine.set_tree(trait_index, left, right, cols)
    counts, p, _ = engine.contingency_fisher(trait_index)
    # ---- counts
    pc = _popcount_rows(bits & mask)
    pct = _popcount_rows(bits & (value & mask))
    want = np.stack([pct, pc - pct, n_pos - pct, (n_mask - n_pos) - (pc - pct)], axis=1)
    assert np.array_equal(counts.astype(np.int64), want)
    tested = (pc > 0) & (pc < n_mask)
    assert np.all((p[tested] >= 0) & (p[tested] <= 1.0))      # planted genes underflow to 0 at N = 5000, as in SciPy
    # ---- exhaustive permutations
    pairs, r, nd = engine.permute(trait_index, P, seed=seed)
    assert np.all(nd == P) and np.all((r >= 0) & (r <= P))
    total, pro, anti = pairs[:, 0].astype(np.int64), pairs[:, 1].astype(np.int64), pairs[:, 2].astype(np.int64)
    assert np.all((pro >= 0) & (anti >= 0) & (pro <= total) & (anti <= total))
    assert np.all(total <= np.minimum(np.minimum(pc, n_mask - pc), min(n_pos, n_mask - n_pos)))
    assert np.all(total[~tested] == 0)
    # ---- the complemented gene matrix
    full = np.zeros(bits.shape[1], dtype=np.uint64)
    for w in range(bits.shape[1]):
        nb = min(64, max(0, N - 64 * w))
        full[w] = np.uint64((1 << nb) - 1) if nb < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    bits_c = bits ^ full
    engine.set_genes(bits_c, N)
    counts_c, p_c, _ = engine.contingency_fisher(trait_index)
    assert np.array_equal(counts_c[:, [2, 3, 0, 1]], counts)
    assert np.array_equal(np.isfinite(p), np.isfinite(p_c))
    both = np.isfinite(p) & np.isfinite(p_c) & (p > 1e-290)
    assert np.max(np.abs(p_c[both] - p[both]) / p[both], initial=0.0) <= FISHER_RTOL
    tiny = np.isfinite(p) & (p <= 1e-290)
    assert np.all(p_c[tiny] <= 2e-290)
    pairs_c, r_c, nd_c = engine.permute(trait_index, P, seed=seed)
    assert np.array_equal(pairs_c[:, [0, 2, 1]], pairs) and np.all(nd_c == P)
    strict = pro != anti
    assert np.array_equal(r_c[strict], r[strict])
    # ---- any subset of genes gives the rows of the full run
    rng = np.random.default_rng(seed)
    idx = np.sort(rng.choice(G, size=min(n_subset, G), replace=False)).astype(np.int64)
    pairs_s, r_s, nd_s = engine.permute(trait_index, P, seed=seed, gene_idx=idx)
    assert np.array_equal(pairs_s, pairs_c[idx]) and np.array_equal(r_s, r_c[idx]) and np.all(nd_s == P)
    assert np.array_equal(engine.pairwise(trait_index, idx), pairs_c[idx])
    # a different seed gives different labellings (some hit count changes) but the same unpermuted walk
    pairs_o, r_o, _ = engine.permute(trait_index, P, seed=seed + 1, gene_idx=idx)
    assert np.array_equal(pairs_o, pairs_s)
    if P >= 10 and len(idx) >= 20:
        assert not np.array_equal(r_o, r_s)
    # ---- the oracle on a few genes at the full N and P (original orientation)
    engine.set_genes(bits, N)
    pick = np.unique(np.concatenate([np.arange(min(2, G)), rng.choice(G, size=min(n_oracle, G), replace=False)]))
    m = synth.unpack_rows(bits[pick], N)
    ref_counts = O.contingency(m, vec)
    assert np.array_equal(counts[pick], ref_counts)
    ref_p = O.fisher(ref_counts)
    ok = np.isfinite(ref_p) & (ref_p > 1e-290)
    assert np.max(np.abs(p[pick][ok] - ref_p[ok]) / ref_p[ok], initial=0.0) <= FISHER_RTOL
    ref = O.permute(left, right, m[:, cols], vec[cols].astype(np.uint8), P=P, seed=seed, trait=trait_index)
    assert np.array_equal(pairs[pick], ref["pairs"]) and np.array_equal(r[pick], ref["r"])
    return {"tested": int(tested.sum()), "strict": int(strict.sum()), "oracle_genes": len(pick)}
