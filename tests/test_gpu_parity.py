"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU
oracle on the same seeded inputs.  Integer results bit-exact; Fisher p within
1e-10 relative of SciPy (the tolerance BASELINE.json's north_star states)."""
import numpy as np
import pytest

from oracle import oracle as O
from scoary_b200 import engine as eng
from scoary_b200 import synth
from scoary_b200 import tree as treemod

pytestmark = pytest.mark.gpu

FISHER_RTOL = 1e-10      # north_star: "within 1e-10 relative for Fisher p-values"


def _dataset(G, N, seed, missing=0.0, T=1):
    traits = synth.make_traits(N, T, seed, missing_frac=missing)
    bits = synth.make_genes_packed(G, N, seed, traits=traits)
    return bits, traits


def _tree_for(N, seed, trait_vec):
    nested = synth.make_tree(N, seed)
    names = synth.isolate_names(N)
    drop = [names[j] for j in range(N) if trait_vec[j] < 0]
    nested = treemod.prune(nested, drop)
    col = {n: j for j, n in enumerate(names)}
    return nested, col


@pytest.mark.parametrize("G,N,missing", [(300, 100, 0.0), (257, 130, 0.05), (2000, 1000, 0.0), (500, 2111, 0.02),
                                         (64, 5000, 0.0), (33, 64, 0.0), (5, 7, 0.0)])
def test_contingency_fisher_hash(engine, G, N, missing):
    bits, traits = _dataset(G, N, 1000 + G + N, missing)
    engine.set_genes(bits, N)
    engine.set_trait_vector(0, traits[0])
    counts, p, h = engine.contingency_fisher(0, want_hash=True)
    m = synth.unpack_rows(bits, N)
    ref_counts = O.contingency(m, traits[0])
    assert np.array_equal(counts, ref_counts)                      # bit-exact
    assert np.array_equal(h, O.pattern_hash(m, traits[0]))         # bit-exact
    ref_p = O.fisher_scipy(ref_counts)                             # the reference's own arithmetic (SciPy)
    ok = ref_p > 1e-290
    rel = np.abs(p - ref_p) / np.maximum(ref_p, 1e-300)
    assert rel[ok].max() <= FISHER_RTOL, (rel[ok].max(), ref_counts[int(np.argmax(np.where(ok, rel, 0)))])
    assert np.all(p[~ok] < 1e-280)
    # the C oracle agrees too (and is what larger cases use)
    rel2 = np.abs(p - O.fisher(ref_counts)) / np.maximum(ref_p, 1e-300)
    assert rel2[ok].max() <= FISHER_RTOL


def test_fisher_edge_tables(engine):
    """all-0 / all-1 genes, genes equal to the trait or its complement, symmetric margins (exact ties)."""
    N = 200
    t = np.zeros(N, dtype=np.int8); t[:100] = 1
    rows = [np.zeros(N), np.ones(N), t == 1, t == 0]
    rng = np.random.default_rng(5)
    for k in range(200):
        r = np.zeros(N); r[rng.choice(N, 100, replace=False)] = 1   # row margin N/2 and col margin N/2 -> ties
        rows.append(r)
    m = np.asarray(rows, dtype=np.uint8)
    engine.set_genes(eng.pack_rows(m), N)
    engine.set_trait_vector(0, t)
    counts, p, _ = engine.contingency_fisher(0)
    ref_counts = O.contingency(m, t)
    assert np.array_equal(counts, ref_counts)
    ref_p = O.fisher_scipy(ref_counts)
    assert p[0] == 1.0 and p[1] == 1.0
    rel = np.abs(p - ref_p) / ref_p
    assert rel.max() <= FISHER_RTOL


@pytest.mark.parametrize("G,N,missing", [(300, 100, 0.0), (200, 257, 0.06), (150, 1000, 0.0), (40, 3000, 0.01), (7, 2, 0.0),
                                         (9, 3, 0.0), (130, 33, 0.0)])
def test_pairwise_bit_exact(engine, G, N, missing):
    bits, traits = _dataset(G, N, 2000 + G + N, missing)
    nested, col = _tree_for(N, 77 + N, traits[0])
    engine.set_genes(bits, N)
    engine.set_trait_vector(0, traits[0])
    names = engine.set_tree_nested(0, nested, col)
    pairs = engine.pairwise(0)
    left, right, onames = O.flatten_tree(nested)
    assert onames == names
    m = synth.unpack_rows(bits, N)
    cols = np.asarray([col[n] for n in names])
    labels = traits[0][cols].astype(np.uint8)
    ref = O.permute(left, right, m[:, cols], labels, P=0)["pairs"]
    assert np.array_equal(pairs, ref)
    # a subset through gene_idx, in scrambled order
    idx = np.random.default_rng(1).permutation(G)[: max(1, G // 3)]
    assert np.array_equal(engine.pairwise(0, idx), ref[idx])


def test_caterpillar_and_balanced_trees(engine):
    """extreme shapes: a comb (stack depth 0) and a perfectly balanced tree (max stack depth)."""
    N = 256
    names = synth.isolate_names(N)
    comb = names[0]
    for n in names[1:]:
        comb = [comb, n]
    level = list(names)
    while len(level) > 1:
        level = [[level[i], level[i + 1]] for i in range(0, len(level), 2)]
    bal = level[0]
    bits, traits = _dataset(200, N, 31337)
    col = {n: j for j, n in enumerate(names)}
    m = synth.unpack_rows(bits, N)
    engine.set_genes(bits, N)
    engine.set_trait_vector(0, traits[0])
    for nested in (comb, bal):
        nm = engine.set_tree_nested(0, nested, col)
        left, right, _ = O.flatten_tree(nested)
        cols = np.asarray([col[n] for n in nm])
        ref = O.permute(left, right, m[:, cols], traits[0][cols].astype(np.uint8), P=0)["pairs"]
        assert np.array_equal(engine.pairwise(0), ref)


@pytest.mark.parametrize("N,P", [(100, 40), (257, 33), (1000, 64)])
def test_shuffles_match_oracle_stream(engine, N, P):
    bits, traits = _dataset(10, N, 99 + N)
    nested, col = _tree_for(N, 5 + N, traits[0])
    engine.set_genes(bits, N)
    engine.set_trait_vector(0, traits[0])
    names = engine.set_tree_nested(0, nested, col)
    cols = np.asarray([col[n] for n in names])
    labels = traits[0][cols].astype(np.uint8)
    got = engine.shuffled_labels(0, P, seed=0xC0FFEE12345, n_leaves=len(names))
    for i in range(P):
        ref = O.shuffle_labels(0xC0FFEE12345, 0, i, labels)
        assert np.array_equal(got[i], ref), i
    assert np.all(got.sum(axis=1) == labels.sum())      # label counts preserved


@pytest.mark.parametrize("G,N,P,missing", [(150, 100, 100, 0.0), (60, 300, 64, 0.04), (40, 1000, 37, 0.0)])
def test_permute_bit_exact(engine, G, N, P, missing):
    bits, traits = _dataset(G, N, 4000 + G + N, missing)
    nested, col = _tree_for(N, 11 + N, traits[0])
    engine.set_genes(bits, N)
    engine.set_trait_vector(1, traits[0])            # a non-zero trait slot: the RNG counter carries it
    names = engine.set_tree_nested(1, nested, col)
    left, right, _ = O.flatten_tree(nested)
    m = synth.unpack_rows(bits, N)
    cols = np.asarray([col[n] for n in names])
    labels = traits[0][cols].astype(np.uint8)
    seed = 20260924
    ref = O.permute(left, right, m[:, cols], labels, P=P, seed=seed, trait=1)
    pairs, r, nd = engine.permute(1, P, seed=seed)
    assert np.array_equal(pairs, ref["pairs"])
    assert np.array_equal(r, ref["r"])
    assert np.all(nd == P)
    # the reference's sequential early-stop rule on the same hit sequence
    ref_es = O.permute(left, right, m[:, cols], labels, P=P, seed=seed, trait=1, early_stop=True)
    _, r2, nd2 = engine.permute(1, P, seed=seed, early_stop=True, rmin=O.rmin_table(P))
    assert np.array_equal(r2, ref_es["r"])
    assert np.array_equal(nd2, ref_es["n_done"])


def test_errors_are_reported(engine):
    from scoary_b200.engine import EngineError
    bits, traits = _dataset(10, 64, 1)
    engine.set_genes(bits, 64)
    with pytest.raises(EngineError):
        engine.contingency_fisher(5)                 # trait 5 never set
    engine.set_trait_vector(0, traits[0])
    with pytest.raises(EngineError):
        engine.pairwise(7)                           # no tree for slot 7
    with pytest.raises(EngineError):                 # not children-before-parents
        engine.set_tree(0, [1, ~0], [~1, ~2], [0, 1, 2])


@pytest.mark.parametrize("G,N,seed", [(300, 17, 1), (2000, 100, 2), (5000, 333, 3), (64, 2, 4), (1000, 65, 5)])
def test_upgma_merge_order_matches_restatement(engine, G, N, seed):
    """GPU tree construction == the reference's merge order (oracle restatement,
    itself pinned on ExampleTree.nwk), including ties: small N and few genes give many equal distances."""
    rng = np.random.default_rng(seed)
    m = (rng.random((G, N)) < rng.uniform(0.05, 0.95, size=(G, 1))).astype(np.uint8)
    if seed == 1:
        m[:, 5] = m[:, 3]          # identical isolates -> distance 0 ties
        m[:, 9] = m[:, 3]
    engine.set_genes(eng.pack_rows(m), N)
    got = engine.upgma()
    assert np.array_equal(got, O.upgma_merges(m))


def test_upgma_example_tree(engine):
    """the reference's golden tree (exampledata/ExampleTree.nwk), byte for byte, from the GPU"""
    import gzip, os
    from scoary_b200 import methods as M
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inputs")
    with gzip.open(os.path.join(gold, "Gene_presence_absence.csv.gz"), "rt") as fh:
        table = M.Csv_to_dic_Roary(fh, ",", [], startcol=14)["Roarydic"]
    engine.set_genes(eng.pack_rows(table.matrix), len(table.strains))
    tree = treemod.from_merges(table.strains, engine.upgma())
    assert treemod.to_scoary_newick(tree) == open(os.path.join(gold, "ExampleTree.nwk")).read().strip()


def test_upgma_matches_reference_golden(engine):
    """GPU tree construction against trees built by the unmodified reference (tests/golden/upgma.json)."""
    import json, os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "upgma.json")
    for u in json.load(open(gold)):
        mat = np.asarray(u["matrix"], dtype=np.uint8)
        names = ["s%d" % k for k in range(mat.shape[1])]
        engine.set_genes(eng.pack_rows(mat), mat.shape[1])
        assert treemod.from_merges(names, engine.upgma()) == u["tree"]


def _comb(names):
    t = names[0]
    for n in names[1:]:
        t = [t, n]
    return t


def _balanced(names):
    level = list(names)
    while len(level) > 1:
        nxt = [[level[i], level[i + 1]] for i in range(0, len(level) - 1, 2)]
        if len(level) % 2:
            nxt.append(level[-1])
        level = nxt
    return level[0]


@pytest.mark.parametrize("N", [126, 127, 128, 129, 130, 254, 255, 256, 257, 383])
def test_walk_mode_boundaries(engine, N):
    """subtree sizes around the 16-bit / 32-bit switch (127 leaves), for three tree shapes"""
    names = synth.isolate_names(N)
    col = {n: j for j, n in enumerate(names)}
    bits, traits = _dataset(70, N, 900 + N)
    m = synth.unpack_rows(bits, N)
    engine.set_genes(bits, N)
    engine.set_trait_vector(0, traits[0])
    for nested in (_comb(names), _balanced(names), synth.make_tree(N, N)):
        order = engine.set_tree_nested(0, nested, col)
        left, right, _ = O.flatten_tree(nested)
        cols = np.asarray([col[n] for n in order])
        ref = O.permute(left, right, m[:, cols], traits[0][cols].astype(np.uint8), P=9, seed=N, trait=0)
        pairs, r, nd = engine.permute(0, 9, seed=N)
        assert np.array_equal(pairs, ref["pairs"]) and np.array_equal(r, ref["r"])


@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 10, 31, 32, 33, 70, 131])
def test_permutation_counts_and_early_stop_boundaries(engine, P):
    G, N = 90, 200
    bits, traits = _dataset(G, N, 4242)
    nested, col = _tree_for(N, 4242, traits[0])
    engine.set_genes(bits, N)
    engine.set_trait_vector(2, traits[0])
    names = engine.set_tree_nested(2, nested, col)
    left, right, _ = O.flatten_tree(nested)
    m = synth.unpack_rows(bits, N)
    cols = np.asarray([col[n] for n in names])
    labels = traits[0][cols].astype(np.uint8)
    idx = np.asarray([5, 5, 0, 89, 17, 3, 44], dtype=np.int64)        # duplicates, arbitrary order, odd count
    for es in (False, True):
        ref = O.permute(left, right, m[np.ix_(idx, cols)], labels, P=P, seed=7, trait=2, early_stop=es)
        pairs, r, nd = engine.permute(2, P, seed=7, gene_idx=idx, early_stop=es, rmin=O.rmin_table(P) if es else None)
        assert np.array_equal(pairs, ref["pairs"])
        assert np.array_equal(r, ref["r"]) and np.array_equal(nd, ref["n_done"]), (P, es)


def test_device_pointer_entry_points(engine):
    """the *_device variants (device buffers, caller's stream) agree with the host-buffer calls"""
    import torch
    G, N, P = 500, 300, 40
    bits, traits = _dataset(G, N, 77)
    nested, col = _tree_for(N, 77, traits[0])
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        d_bits = torch.from_numpy(bits.view(np.int64)).to(dev)
        engine.set_stream(stream.cuda_stream)
        try:
            engine.set_genes_device(d_bits.data_ptr(), G, N, bits.shape[1])
            engine.set_trait_vector(0, traits[0])
            engine.set_tree_nested(0, nested, col)
            d_counts = torch.empty((G, 4), dtype=torch.int32, device=dev)
            d_p = torch.empty(G, dtype=torch.float64, device=dev)
            d_pairs = torch.empty((G, 3), dtype=torch.int32, device=dev)
            d_r = torch.empty(G, dtype=torch.int32, device=dev)
            d_nd = torch.empty(G, dtype=torch.int32, device=dev)
            engine.contingency_fisher_device(0, d_counts.data_ptr(), d_p.data_ptr())
            engine.permute_device(0, G, P, 5, d_pairs.data_ptr(), d_r.data_ptr(), d_nd.data_ptr())
            stream.synchronize()
            got = (d_counts.cpu().numpy(), d_p.cpu().numpy(), d_pairs.cpu().numpy(), d_r.cpu().numpy())
        finally:
            engine.set_stream(0)
    engine.set_genes(bits, N)
    engine.set_trait_vector(0, traits[0])
    engine.set_tree_nested(0, nested, col)
    counts, p, _ = engine.contingency_fisher(0)
    pairs, r, nd = engine.permute(0, P, seed=5)
    assert np.array_equal(got[0], counts) and np.array_equal(got[1], p)
    assert np.array_equal(got[2], pairs) and np.array_equal(got[3], r)


def test_large_isolate_count_paths(engine):
    """N = 20 000 isolates: the log-factorial LUT no longer fits shared memory (global-memory LUT
    path of the Fisher kernel), label vectors take 628 words (few permutations per launch), and a
    comb tree is 19 999 levels deep (pure caterpillar program, widened to 32-bit keys after 127 leaves)."""
    N, G, P = 20000, 48, 6
    bits, traits = _dataset(G, N, 31415, 0.01)
    names = synth.isolate_names(N)
    col = {n: j for j, n in enumerate(names)}
    keep = [n for j, n in enumerate(names) if traits[0][j] >= 0]
    nested = _comb(keep)
    m = synth.unpack_rows(bits, N)
    engine.set_genes(bits, N)
    engine.set_trait_vector(0, traits[0])
    counts, p, _ = engine.contingency_fisher(0)
    ref_counts = O.contingency(m, traits[0])
    assert np.array_equal(counts, ref_counts)
    ref_p = O.fisher(ref_counts)
    ok = ref_p > 1e-290
    assert np.max(np.abs(p - ref_p)[ok] / ref_p[ok]) <= FISHER_RTOL
    order = engine.set_tree_nested(0, nested, col)
    left, right, _ = O.flatten_tree(nested)
    cols = np.asarray([col[n] for n in order])
    ref = O.permute(left, right, m[:, cols], traits[0][cols].astype(np.uint8), P=P, seed=3, trait=0)
    pairs, r, nd = engine.permute(0, P, seed=3)
    assert np.array_equal(pairs, ref["pairs"]) and np.array_equal(r, ref["r"])


def test_empty_selections(engine):
    bits, traits = _dataset(20, 64, 5)
    nested, col = _tree_for(64, 5, traits[0])
    engine.set_genes(bits, 64)
    engine.set_trait_vector(0, traits[0])
    engine.set_tree_nested(0, nested, col)
    none = np.zeros(0, dtype=np.int64)
    assert engine.pairwise(0, none).shape == (0, 3)
    pairs, r, nd = engine.permute(0, 10, gene_idx=none)
    assert pairs.shape == (0, 3) and r.shape == (0,) and nd.shape == (0,)


@pytest.mark.parametrize("G,N,P,S", [(400, 300, 700, 3), (400, 300, 1100, 41), (200, 1000, 600, 130), (64, 5000, 1500, 64),
                                     (50, 129, 513, 1)])
def test_transposed_launches_match_oracle_and_default_shape(engine, G, N, P, S):
    """Few genes x many permutations (what is left after decideifbreak, methods.py:1022-1024, :1295-1310): K5 runs
    with threads = labellings (sb_set_permute_mode 2).  Same r / n_done as the oracle's Permute and as the
    threads = genes shape, exhaustive and with the reference's early stop; the automatic choice picks it here."""
    bits, traits = _dataset(G, N, 6000 + G + N, 0.02)
    nested, col = _tree_for(N, 21 + N, traits[0])
    engine.set_genes(bits, N)
    engine.set_trait_vector(3, traits[0])
    names = engine.set_tree_nested(3, nested, col)
    left, right, _ = O.flatten_tree(nested)
    m = synth.unpack_rows(bits, N)
    cols = np.asarray([col[n] for n in names])
    labels = traits[0][cols].astype(np.uint8)
    idx = np.random.default_rng(S).choice(G, size=S, replace=False).astype(np.int64)
    idx[0] = 0                                                    # a planted gene: never stops early
    rmin = O.rmin_table(P)
    try:
        for es in (False, True):
            ref = O.permute(left, right, m[np.ix_(idx, cols)], labels, P=P, seed=99, trait=3, early_stop=es)
            out = {}
            for mode in (1, 2, 0):
                engine.set_permute_mode(mode)
                engine.stats_reset()
                pairs, r, nd = engine.permute(3, P, seed=99, gene_idx=idx, early_stop=es, rmin=rmin if es else None)
                out[mode] = engine.stats()["calls_transposed"]
                assert np.array_equal(pairs, ref["pairs"]), (mode, es)
                assert np.array_equal(r, ref["r"]) and np.array_equal(nd, ref["n_done"]), (mode, es)
            assert out[1] == 0 and out[2] == 1
            if not es and N >= 5000:
                assert out[0] == 1                                # auto: threads = labellings fills the GPU better here
    finally:
        engine.set_permute_mode(0)


@pytest.mark.parametrize("G,N,T,missing", [(300, 100, 2, 0.0), (1000, 1000, 4, 0.03), (257, 5000, 3, 0.01), (64, 333, 11, 0.05)])
def test_multi_trait_fisher_pass_equals_single_calls(engine, G, N, T, missing):
    """sb_contingency_fisher_multi: every gene row read once for all traits (the reference loops traits outside
    genes, methods.py:771, :791).  Bit-identical to T single-trait calls -- counts, hashes and p -- and the counts
    equal the oracle's; T = 11 exceeds the 8 traits one launch stages (two launches)."""
    bits, traits = _dataset(G, N, 7000 + G + N, missing, T=T)
    engine.set_genes(bits, N)
    for t in range(T):
        engine.set_trait_vector(5 + t, traits[t])                 # slots 5 ..: consecutive slots, not from 0
    c_all, p_all, h_all = engine.contingency_fisher_multi(5, T, want_hash=True)
    m = synth.unpack_rows(bits, N)
    for t in range(T):
        c1, p1, h1 = engine.contingency_fisher(5 + t, want_hash=True)
        assert np.array_equal(c_all[t], c1) and np.array_equal(h_all[t], h1)
        assert np.array_equal(p_all[t].view(np.uint64), p1.view(np.uint64))
        ref_counts = O.contingency(m, traits[t])
        assert np.array_equal(c1, ref_counts)
        ref_p = O.fisher(ref_counts)
        ok = ref_p > 1e-290
        assert np.max(np.abs(p1 - ref_p)[ok] / ref_p[ok]) <= FISHER_RTOL


def test_fisher_near_mode_and_tail_paths(engine):
    """The Fisher kernel sums the complement for tables near the mode (p >= 0.01) and the two tails otherwise:
    6 000 tables built to straddle the switch (|z| up to 3.5 at N = 5 000), against SciPy itself."""
    N, G = 5000, 6000
    rng = np.random.default_rng(123)
    t = (rng.random(N) < 0.35).astype(np.int8)
    n_pos = int(t.sum())
    rows = np.zeros((G, N), dtype=np.uint8)
    pos, neg = np.flatnonzero(t == 1), np.flatnonzero(t == 0)
    for g in range(G):
        r1 = int(rng.integers(100, N - 100))
        mean = r1 * n_pos / N
        sd = (mean * (1 - r1 / N) * (1 - n_pos / N)) ** 0.5
        a = int(np.clip(round(mean + rng.uniform(-3.5, 3.5) * sd), max(0, r1 - len(neg)), min(r1, n_pos)))
        rows[g, rng.choice(pos, a, replace=False)] = 1
        rows[g, rng.choice(neg, r1 - a, replace=False)] = 1
    engine.set_genes(eng.pack_rows(rows), N)
    engine.set_trait_vector(0, t)
    counts, p, _ = engine.contingency_fisher(0)
    ref_counts = O.contingency(rows, t)
    assert np.array_equal(counts, ref_counts)
    ref_p = O.fisher(ref_counts)
    rel = np.abs(p - ref_p) / ref_p
    assert rel.max() <= FISHER_RTOL, rel.max()
    pick = np.argsort(np.abs(ref_p - 0.01))[:300]                   # around the switch: SciPy's own arithmetic
    sp = O.fisher_scipy(ref_counts[pick])
    assert np.max(np.abs(p[pick] - sp) / sp) <= FISHER_RTOL


@pytest.mark.parametrize("shape", ["random", "balanced"])
def test_maximum_tree_size(engine, shape):
    """32 766 isolates, the limit include/scoary_b200.h states, for the two shapes with the longest programs
    (a random-join tree compiles to ~16 400 ops, a balanced one to ~21 400 -- both beyond the 12 288
    ops round 1 could hold).  Pair counts and permutation hit counts against the oracle."""
    N, G, P = 32766, 12, 5
    bits, traits = _dataset(G, N, 27182)
    names = synth.isolate_names(N)
    col = {n: j for j, n in enumerate(names)}
    nested = synth.make_tree(N, 99) if shape == "random" else _balanced(names)
    m = synth.unpack_rows(bits, N)
    engine.set_genes(bits, N)
    engine.set_trait_vector(0, traits[0])
    order = engine.set_tree_nested(0, nested, col)
    left, right, _ = O.flatten_tree(nested)
    cols = np.asarray([col[n] for n in order])
    ref = O.permute(left, right, m[:, cols], traits[0][cols].astype(np.uint8), P=P, seed=4, trait=0)
    pairs, r, nd = engine.permute(0, P, seed=4)
    assert np.array_equal(pairs, ref["pairs"]) and np.array_equal(r, ref["r"]) and np.all(nd == P)
    from scoary_b200.engine import EngineError
    with pytest.raises(EngineError):                               # one leaf more is refused by sb_set_tree itself
        big = synth.isolate_names(N + 1)
        engine.set_genes(np.zeros((2, eng.words_for(N + 1)), dtype=np.uint64), N + 1)
        engine.set_trait_vector(0, np.zeros(N + 1, dtype=np.int8))
        engine.set_tree_nested(0, _comb(big), {n: j for j, n in enumerate(big)})


@pytest.mark.parametrize("n,ties", [(1, 0), (2, 1), (1000, 0), (5000, 40), (200_000, 500), (1_000_003, 2000)])
def test_device_epilogue_equals_host_implementation(engine, n, ties):
    """Bonferroni, Benjamini-Hochberg with the reference's tie rule (methods.py:903-919) and the
    stable p-sort (:1448-1454) on the device -- bit-identical to the host code the golden result files pin, including
    runs of tied p-values, untested genes and a number_of_tests that differs from the row count (--collapse)."""
    from scoary_b200 import methods as M
    rng = np.random.default_rng(n)
    p = rng.random(n) ** 3
    p[rng.random(n) < 0.02] = 1.0
    for _ in range(ties):                                           # tie runs of 2 .. 6 genes
        src = int(rng.integers(n))
        p[rng.integers(n, size=int(rng.integers(1, 6)))] = p[src]
    keep = (rng.random(n) < 0.97)
    if n <= 2:
        keep[:] = True
    for n_tests in (0, max(1, int(keep.sum()) - 3)):
        order, bonf, bh = engine.adjust_pvalues(p, keep, n_tests)
        idx = np.flatnonzero(keep)
        m = n_tests or len(idx)
        want_order = idx[np.argsort(p[idx], kind="stable")]
        assert np.array_equal(order, want_order)
        want_bh = np.minimum(M.benjamini_hochberg(p[want_order], m), 1.0)
        assert np.array_equal(bh[want_order].view(np.uint64), want_bh.view(np.uint64))
        assert np.array_equal(bonf[idx].view(np.uint64), np.minimum(p[idx] * m, 1.0).view(np.uint64))
        assert np.all(np.isnan(bh[~keep])) and np.all(np.isnan(bonf[~keep]))


def test_device_binomial_test_matches_scipy(engine):
    """ss.binom_test(k, n, 0.5) (methods.py:1267-1275) on the device against SciPy's binomtest, every k for small n and
    random (k, n) up to the 16 383 pairs a 32 766-leaf tree can hold."""
    from scipy import stats as ss
    ks, ns = [], []
    for n in range(0, 40):
        for k in range(0, n + 1):
            ks.append(k); ns.append(n)
    rng = np.random.default_rng(3)
    for n in rng.integers(40, 16384, size=3000):
        n = int(n)
        k = int(np.clip(round(n / 2 + rng.normal() * 3 * (n ** 0.5) / 2), 0, n)) if rng.random() < 0.8 else int(rng.integers(0, n + 1))
        ks.append(k); ns.append(n)
    k, n = np.asarray(ks), np.asarray(ns)
    got = engine.binom_two_sided(k, n)
    assert np.all(np.isnan(got[n == 0]))
    ok = n > 0
    want = np.asarray(ss.binomtest(k[ok], n[ok], 0.5).pvalue, dtype=np.float64)
    big = want > 1e-290
    rel = np.abs(got[ok][big] - want[big]) / want[big]
    assert rel.max() <= 1e-11, rel.max()
    # SciPy itself degrades down there: for n > 1074 its two-sided sum underflows to exactly 0 where the true value
    # (and the device's direct sum) is still a normal number, e.g. k = 29, n = 1077: 8.4e-268
    tiny = ~(got[ok][~big] <= 1e-200)
    assert not tiny.any(), (k[ok][~big][tiny][:5], n[ok][~big][tiny][:5], got[ok][~big][tiny][:5], want[~big][tiny][:5])


def test_reference_rule_mode_many_rounds(engine):
    """Reference-rule mode at the bench's shape in miniature: thousands of permutations, a tree large enough that the
    constant pool holds few labellings, so the survivors run through > 100 device-paced rounds (each with its own
    list-length counter).  r and n_done equal the oracle's sequential Permute."""
    G, N, P = 700, 3000, 4000
    bits, traits = _dataset(G, N, 515)
    nested, col = _tree_for(N, 516, traits[0])
    engine.set_genes(bits, N)
    engine.set_trait_vector(0, traits[0])
    names = engine.set_tree_nested(0, nested, col)
    left, right, _ = O.flatten_tree(nested)
    m = synth.unpack_rows(bits, N)
    cols = np.asarray([col[n] for n in names])
    labels = traits[0][cols].astype(np.uint8)
    engine.set_permute_mode(1)
    try:
        engine.stats_reset()
        pairs, r, nd = engine.permute(0, P, seed=31, early_stop=True, rmin=O.rmin_table(P))
        walks = engine.stats()["tests_walks"]
    finally:
        engine.set_permute_mode(0)
    pick = np.concatenate([np.arange(4), np.flatnonzero(nd == P)[:4], np.flatnonzero(nd < P)[:10]])
    ref = O.permute(left, right, m[np.ix_(pick, cols)], labels, P=P, seed=31, trait=0, early_stop=True)
    assert np.array_equal(pairs[pick], ref["pairs"])
    assert np.array_equal(r[pick], ref["r"]) and np.array_equal(nd[pick], ref["n_done"])
    assert (nd == P).sum() >= 5 and (nd == 31).sum() > G // 2
    assert G + int(nd.sum()) <= walks <= G + int(nd.sum()) + 64 * G      # walks really done: n_done + slice overshoot


@pytest.mark.parametrize("G,N,P,parts", [(300, 200, 100, 3), (64, 1000, 257, 8), (700, 129, 64, 2)])
def test_permutation_ranges_add_up_to_the_job(engine, G, N, P, parts):
    """sb_permute_range: the labelling of permutation i depends on (seed, trait, i) alone, so disjoint ranges of the
    permutations -- the way N GPUs split a job whose gene shards would be too small -- give hit counts that add up to
    sb_permute's r, which equals the oracle's."""
    from scoary_b200 import distributed as D
    bits, traits = _dataset(G, N, 8000 + G + N, 0.01)
    nested, col = _tree_for(N, 31 + N, traits[0])
    engine.set_genes(bits, N)
    engine.set_trait_vector(2, traits[0])
    names = engine.set_tree_nested(2, nested, col)
    left, right, _ = O.flatten_tree(nested)
    m = synth.unpack_rows(bits, N)
    cols = np.asarray([col[n] for n in names])
    ref = O.permute(left, right, m[:, cols], traits[0][cols].astype(np.uint8), P=P, seed=77, trait=2)
    pairs, r, nd = engine.permute(2, P, seed=77)
    assert np.array_equal(r, ref["r"]) and np.array_equal(pairs, ref["pairs"])
    total = np.zeros(G, dtype=np.int64)
    for rank in range(parts):
        first, count = D.permutation_range(P, parts, rank)
        pr, rr = engine.permute_range(2, first, count, seed=77)
        assert np.array_equal(pr, pairs)
        total += rr
    assert np.array_equal(total, r)
    idx = np.asarray([3, 0, 17], dtype=np.int64)
    _, r_sub = engine.permute_range(2, 5, P - 5, seed=77, gene_idx=idx)
    _, r_head = engine.permute_range(2, 0, 5, seed=77, gene_idx=idx)
    assert np.array_equal(r_sub + r_head, r[idx])
