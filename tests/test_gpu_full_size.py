"""BASELINE.json's sizes on the GPU.  The oracle cannot redo 5e7 walks, so the whole result is
checked through size-independent properties (tests/full_size_properties.py) and a handful of genes
are compared with the oracle directly at the full isolate and permutation counts."""
import pytest

import full_size_properties as props

pytestmark = pytest.mark.gpu


def test_c3_shape_full_size(engine):
    """configs C3: 50 000 genes x 5 000 isolates x 1 trait, 1 000 permutations (the bench workload)."""
    out = props.check(engine, 50000, 5000, 1000, seed=20260903, n_oracle=6, n_subset=1000)
    assert out["tested"] > 49000


def test_c4_isolate_count_with_missing_values(engine):
    """C4's isolate count (10 000) with 2 % missing trait values (pruned tree, masked counts); fewer genes
    and permutations than C4 itself (4e9 walks are a multi-GPU job)."""
    out = props.check(engine, 4096, 10000, 64, seed=20260904, missing=0.02, n_oracle=4, n_subset=500)
    assert out["tested"] > 4000


def test_c5_shape_many_variants(engine):
    """C5's shape: many variants, 2 000 isolates (gene count scaled to 200 000 to keep the test short)."""
    out = props.check(engine, 200000, 2000, 100, seed=20260905, n_oracle=4, n_subset=1000)
    assert out["tested"] > 190000


def test_north_star_10000_permutations(engine):
    """BASELINE.json north_star: 50 000 genes x 5 000 isolates x 10 000 permutations (Permute, methods.py:1314-1369).
    110 K5 launches per pass instead of 11; four oracle genes redo all 10 000 labellings each."""
    out = props.check(engine, 50000, 5000, 10000, seed=20260903, n_oracle=2, n_subset=600)
    assert out["tested"] > 49000 and out["oracle_genes"] >= 4


def test_c4_per_gpu_shard_four_traits(engine):
    """configs C4, the shard one of eight GPUs owns: 12 500 genes x 10 000 isolates x FOUR traits (2 % missing values:
    each trait has its own mask, pruned tree and labellings) x 1 000 permutations; the four Fisher passes run as
    one multi-trait launch (sb_contingency_fisher_multi) and must equal the single-trait calls bit for bit."""
    out = props.check(engine, 12500, 10000, 1000, seed=20260904, missing=0.02, n_oracle=2, n_subset=500, T=4)
    assert out["tested"] > 4 * 12000


def test_c5_per_gpu_shard(engine):
    """configs C5, the shard one of eight GPUs owns: 125 000 variants x 2 000 isolates x 1 000 permutations."""
    out = props.check(engine, 125000, 2000, 1000, seed=20260905, n_oracle=4, n_subset=1000)
    assert out["tested"] > 120000
