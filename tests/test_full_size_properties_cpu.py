"""The full-size property checker (tests/full_size_properties.py) run against the oracle-backed
FakeEngine at a small size: proves the properties hold for the reference's algorithm and that
the checker is sound before it is pointed at the GPU at BASELINE.json's sizes."""
import pytest

from fake_engine import FakeEngine
import full_size_properties as props


@pytest.mark.parametrize("G,N,P,missing,T", [(120, 70, 12, 0.0, 1), (90, 131, 10, 0.05, 1), (60, 50, 10, 0.05, 3)])
def test_properties_hold_for_the_oracle(G, N, P, missing, T):
    out = props.check(FakeEngine(), G, N, P, seed=11, missing=missing, n_oracle=5, n_subset=40, T=T)
    assert out["tested"] > 0 and out["strict"] > 0
