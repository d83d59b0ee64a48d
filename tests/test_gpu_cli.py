"""Drop-in check on the GPU: the real CLI (scoary_b200.methods.main -> libscoary_b200.so)
on the reference's example data against the reference's own result files.  Integer columns
and row order exact; Fisher-derived columns within 1e-10 relative (north_star tolerance)."""
import csv
import gzip
import os
import shutil

import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INT_COLS = ["Number_pos_present_in", "Number_neg_present_in", "Number_pos_not_present_in", "Number_neg_not_present_in",
            "Max_Pairwise_comparisons", "Max_supporting_pairs", "Max_opposing_pairs"]
FLOAT_COLS = ["Sensitivity", "Specificity", "Odds_ratio", "Naive_p", "Bonferroni_p", "Benjamini_H_p",
              "Best_pairwise_comp_p", "Worst_pairwise_comp_p"]
RTOL = 1e-10


@pytest.fixture()
def inputs(tmp_path):
    g = tmp_path / "Gene_presence_absence.csv"
    with gzip.open(os.path.join(GOLD, "inputs", "Gene_presence_absence.csv.gz"), "rb") as fi, open(g, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    return {"g": str(g), "t": os.path.join(GOLD, "inputs", "Tetracycline_resistance.csv"),
            "r": os.path.join(GOLD, "inputs", "Restrict_to.csv"), "out": str(tmp_path / "out")}


def _rows(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as fh:
        rows = list(csv.reader(fh))
    return rows[0], rows[1:]


def _run(argv):
    from scoary_b200 import methods as M
    M._ENGINE = None            # a real engine: Engine() raises without the .so or a GPU
    with pytest.raises(SystemExit) as ex:
        M.main(argv=argv)
    assert ex.value.code == 0
    from scoary_b200.engine import Engine
    assert isinstance(M._ENGINE, Engine)


def _compare(got_path, gold_path, key_cols=1):
    hdr, got = _rows(got_path)
    ghdr, gold = _rows(gold_path)
    assert hdr == ghdr
    assert len(got) == len(gold)
    col = {h: i for i, h in enumerate(hdr)}
    assert sorted(tuple(r[:key_cols]) for r in got) == sorted(tuple(r[:key_cols]) for r in gold)
    by_gene = {tuple(r[:key_cols]): r for r in got}
    assert len(by_gene) == len(got)
    for g in gold:
        r = by_gene[tuple(g[:key_cols])]
        assert r[1:3] == g[1:3]
        for c in INT_COLS:
            if c in col:
                assert r[col[c]] == g[col[c]], (g[0], c)
        for c in FLOAT_COLS:
            if c in col:
                a, b = float(r[col[c]]), float(g[col[c]])
                if b > 1e-290:
                    assert a == b or abs(a - b) <= RTOL * abs(b), (g[0], c, a, b)
    # Row order: both files are sorted by naive p.  SciPy's own p differs by an ulp between
    # symmetric variants of one table (row/column swap, transpose), so rows whose p agree to
    # 1e-10 may come out in a different order; any row that moved must be such a near-tie.
    ip = col["Naive_p"]
    swapped = 0
    for a, b in zip(got, gold):
        pa, pb = float(a[ip]), float(b[ip])
        assert pa == pb or abs(pa - pb) <= RTOL * abs(pb), ("sorted p sequence differs", a[0], b[0], pa, pb)
        swapped += a[:key_cols] != b[:key_cols]
    return swapped, len(gold)


@pytest.mark.parametrize("name,extra", [("default", []), ("nopairwise", ["--no_pairwise"]),
                                        ("all", ["-p", "1.0", "-c", "I"]),
                                        ("collapse", ["-p", "1.0", "-c", "I", "--collapse"]),
                                        ("advanced", ["-p", "0.01", "1E-5", "-c", "B", "EPW", "--collapse", "-m", "50", "-u"])])
def test_cli_matches_reference_results(name, extra, inputs):
    _run(["-g", inputs["g"], "-t", inputs["t"], "-o", inputs["out"], "--no-time"] + extra)
    for trait in ("Tetracycline_resistance", "Bogus_trait"):
        gold = os.path.join(GOLD, name, trait + ".results.csv")
        gold = gold if os.path.exists(gold) else gold + ".gz"
        _compare(os.path.join(inputs["out"], trait + ".results.csv"), gold)


def test_cli_restricted(inputs):
    _run(["-g", inputs["g"], "-t", inputs["t"], "-o", inputs["out"], "--no-time", "-r", inputs["r"], "-p", "1.0"])
    for trait in ("Tetracycline_resistance", "Bogus_trait"):
        _compare(os.path.join(inputs["out"], trait + ".results.csv"),
                 os.path.join(GOLD, "restrict", trait + ".results.csv.gz"))


def test_cli_permutations_statistically_like_reference(inputs):
    """-e 200: the reference's Empirical_p comes from an unseeded Mersenne Twister, ours from
    Philox; both estimate the same tail probability with the same early-stop rule, so they must
    agree within binomial sampling error.  Everything else in the file is deterministic."""
    _run(["-g", inputs["g"], "-t", inputs["t"], "-o", inputs["out"], "--no-time", "-e", "200", "-c", "I", "EPW",
          "-p", "0.05", "0.05"])
    hdr, got = _rows(os.path.join(inputs["out"], "Tetracycline_resistance.results.csv"))
    ghdr, gold = _rows(os.path.join(GOLD, "perm", "Tetracycline_resistance.results.csv"))
    assert hdr == ghdr and sorted(r[0] for r in got) == sorted(r[0] for r in gold)
    ie = hdr.index("Empirical_p")
    by_gene = {r[0]: r for r in got}
    for g in gold:
        r = by_gene[g[0]]
        assert r[:ie][3:7] == g[:ie][3:7] and r[13:16] == g[13:16]
        a, b = float(r[ie]), float(g[ie])
        n = 200
        se = (max(b, 1.0 / n) * (1 - min(b, 0.99)) / 30) ** 0.5     # >= 31 permutations always run
        assert abs(a - b) <= 6 * se + 2.0 / 32, (r[0], a, b)


def test_permute_and_walk_mirror_functions():
    """ConvertUPGMAtoPhyloTree / Permute with the reference's argument types, on the GPU."""
    import json
    from scoary_b200 import methods as M
    M._ENGINE = None
    walks = json.load(open(os.path.join(GOLD, "walks.json")))
    for w in walks[:40]:
        out = M.ConvertUPGMAtoPhyloTree(w["tree"], w["gtc"])
        assert [out["Total"], out["Pro"], out["Anti"]] == w["out"]
    w = next(x for x in walks if len(x["gtc"]) >= 64 and x["out"][0] > 3)
    emp = M.Permute(w["tree"], w["gtc"], 100, {"I": 0.05})
    assert 0.0 < emp <= 1.0

