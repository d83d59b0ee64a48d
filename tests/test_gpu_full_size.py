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
