"""A few seeds of the differential fuzzer (tests/diff_fuzz_reference.py): random small inputs and command
lines through the unmodified reference CLI and through this repo's CLI; every output file must be
byte-identical.  Needs /root/reference, so it runs in the build container only."""
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")


@pytest.mark.parametrize("seed", list(range(2000, 2012)))      # CSV cases and (about one in five) VCF cases
def test_same_files_as_the_reference(seed):
    import diff_fuzz_reference as F
    argv, problems, _ = F.one(seed)
    assert not problems, (problems, " ".join(argv[4:]))
