"""Drop-in proof at the reference's own call sites: the UNMODIFIED reference main()
(scoary/methods.py:49-330, imported through oracle/ref_shim.py) parses the files, builds its tree
and then calls Setup_results (:278) and StoreResults (:287) -- which are swapped for this repo's
functions of the same name.  The result files must be byte-identical to the reference's own.
Runs only where /root/reference exists (the build container); the GPU part is the oracle-backed
FakeEngine here, and tests/test_gpu_cli.py repeats the comparison with the real engine."""
import gzip
import os
import shutil
import sys

import pytest

from oracle import ref_shim
from scoary_b200 import methods as M
from fake_engine import FakeEngine

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")


@pytest.mark.parametrize("name,extra", [("default", []), ("collapse", ["-p", "1.0", "-c", "I", "--collapse"]),
                                        ("advanced", ["-p", "0.01", "1E-5", "-c", "B", "EPW", "--collapse", "-m", "50", "-u"])])
def test_reference_main_with_our_functions(name, extra, tmp_path, monkeypatch):
    ref = ref_shim.load()
    monkeypatch.setattr(M, "_ENGINE", FakeEngine())
    monkeypatch.setattr(ref, "Setup_results", M.Setup_results)      # the call site at methods.py:278
    monkeypatch.setattr(ref, "StoreResults", M.StoreResults)        # the call site at methods.py:287
    g = tmp_path / "Gene_presence_absence.csv"
    with gzip.open(os.path.join(GOLD, "inputs", "Gene_presence_absence.csv.gz"), "rb") as fi, open(g, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    out = str(tmp_path / "out")
    argv = ["-g", str(g), "-t", os.path.join(GOLD, "inputs", "Tetracycline_resistance.csv"), "-o", out, "--no-time"] + extra
    monkeypatch.setattr(sys, "argv", ["scoary"] + argv)
    with pytest.raises(SystemExit) as ex:
        ref.main()
    assert ex.value.code == 0
    for trait in ("Tetracycline_resistance", "Bogus_trait"):
        gold = os.path.join(GOLD, name, trait + ".results.csv")
        op = open
        if not os.path.exists(gold):
            gold, op = gold + ".gz", gzip.open
        with op(gold, "rt") as fh:
            want = fh.read()
        assert open(os.path.join(out, trait + ".results.csv")).read() == want, (name, trait)


@pytest.mark.parametrize("name,extra", [("default", []), ("all", ["-p", "1.0", "-c", "I"])])
def test_reference_main_with_our_pairwise_comparisons(name, extra, tmp_path, monkeypatch):
    """The other seam of SURVEY 8(b): the reference's own Setup_results and StoreTraitResult, with
    PairWiseComparisons((domain, argdict)) (call site methods.py:1105, def :1208) swapped for this repo's.
    Its GTC argument is then the reference's dict of "AB"/"Ab"/"aB"/"ab" strings."""
    ref = ref_shim.load()
    monkeypatch.setattr(M, "_ENGINE", FakeEngine())
    monkeypatch.setattr(ref, "PairWiseComparisons", M.PairWiseComparisons)
    g = tmp_path / "Gene_presence_absence.csv"
    with gzip.open(os.path.join(GOLD, "inputs", "Gene_presence_absence.csv.gz"), "rb") as fi, open(g, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    out = str(tmp_path / "out")
    argv = ["-g", str(g), "-t", os.path.join(GOLD, "inputs", "Tetracycline_resistance.csv"), "-o", out, "--no-time"] + extra
    monkeypatch.setattr(sys, "argv", ["scoary"] + argv)
    with pytest.raises(SystemExit) as ex:
        ref.main()
    assert ex.value.code == 0
    for trait in ("Tetracycline_resistance", "Bogus_trait"):
        gold = os.path.join(GOLD, name, trait + ".results.csv")
        op = open
        if not os.path.exists(gold):
            gold, op = gold + ".gz", gzip.open
        with op(gold, "rt") as fh:
            want = fh.read()
        assert open(os.path.join(out, trait + ".results.csv")).read() == want, (name, trait)
