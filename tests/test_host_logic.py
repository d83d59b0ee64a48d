"""Host logic of the reference-interface mirror (parsing, UPGMA, collapse, Bonferroni/BH,
sorting, filtering, CSV text) checked against the reference's golden result files, with the
GPU part replaced by the oracle-backed FakeEngine (tests/fake_engine.py).  No GPU needed."""
import gzip
import os
import shutil

import numpy as np
import pytest

from scoary_b200 import methods as M
from scoary_b200 import tree as treemod
from fake_engine import FakeEngine

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture()
def inputs(tmp_path):
    g = tmp_path / "Gene_presence_absence.csv"
    with gzip.open(os.path.join(GOLD, "inputs", "Gene_presence_absence.csv.gz"), "rb") as fi, open(g, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    return {"g": str(g), "t": os.path.join(GOLD, "inputs", "Tetracycline_resistance.csv"),
            "r": os.path.join(GOLD, "inputs", "Restrict_to.csv"), "n": os.path.join(GOLD, "inputs", "ExampleTree.nwk"),
            "out": str(tmp_path / "out")}


@pytest.fixture()
def fake_engine(monkeypatch):
    fe = FakeEngine()
    monkeypatch.setattr(M, "_ENGINE", fe)
    return fe


def _read(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as fh:
        return fh.read()


def _run(argv):
    with pytest.raises(SystemExit) as ex:
        M.main(argv=argv)
    assert ex.value.code == 0


SCENARIOS = {
    "default": [],
    "nopairwise": ["--no_pairwise"],
    "advanced": ["-p", "0.01", "1E-5", "-c", "B", "EPW", "--collapse", "-m", "50", "-u"],
    "all": ["-p", "1.0", "-c", "I"],
    "collapse": ["-p", "1.0", "-c", "I", "--collapse"],
}


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_cli_text_identical_to_reference(name, inputs, fake_engine):
    """Whole result files, byte for byte (header, quoting, row order, float text)."""
    _run(["-g", inputs["g"], "-t", inputs["t"], "-o", inputs["out"], "--no-time"] + SCENARIOS[name])
    for trait in ("Tetracycline_resistance", "Bogus_trait"):
        gold = os.path.join(GOLD, name, trait + ".results.csv")
        gold = gold if os.path.exists(gold) else gold + ".gz"
        assert _read(os.path.join(inputs["out"], trait + ".results.csv")) == _read(gold), (name, trait)
    if name == "advanced":
        assert _read(os.path.join(inputs["out"], "Tree.nwk")) == _read(os.path.join(GOLD, "advanced", "Tree.nwk"))


def test_cli_restrict_and_custom_tree(inputs, fake_engine):
    _run(["-g", inputs["g"], "-t", inputs["t"], "-o", inputs["out"], "--no-time", "-r", inputs["r"], "-p", "1.0"])
    for trait in ("Tetracycline_resistance", "Bogus_trait"):
        assert _read(os.path.join(inputs["out"], trait + ".results.csv")) == \
            _read(os.path.join(GOLD, "restrict", trait + ".results.csv.gz"))
    # -n with the reference's own tree file gives the same result as the internal UPGMA tree
    out2 = inputs["out"] + "2"
    _run(["-g", inputs["g"], "-t", inputs["t"], "-o", out2, "--no-time", "-n", inputs["n"]])
    assert _read(os.path.join(out2, "Tetracycline_resistance.results.csv")) == \
        _read(os.path.join(GOLD, "default", "Tetracycline_resistance.results.csv"))


def test_first_row_is_the_reference_ci_golden(inputs, fake_engine):
    """tests/test_scoary_output.py:12-14 of the reference, same tolerances."""
    import csv
    import json
    ref = json.load(open(os.path.join(GOLD, "tetrcg_first_row.json")))["row"]
    _run(["-g", inputs["g"], "-t", inputs["t"], "-o", inputs["out"], "--no-time"])
    with open(os.path.join(inputs["out"], "Tetracycline_resistance.results.csv")) as fh:
        rows = list(csv.reader(fh))
    d = rows[1]
    assert d[0:3] == ref[0:3]
    assert [int(x) for x in d[3:7]] == ref[3:7]
    assert abs(float(d[7]) - ref[7]) <= 0.01 and abs(float(d[8]) - ref[8]) <= 0.01 and abs(float(d[9]) - ref[9]) <= 0.1
    assert abs(float(d[10]) - ref[10]) <= 1e-15 and abs(float(d[11]) - ref[11]) <= 1e-12
    assert abs(float(d[12]) - ref[12]) <= 1e-12
    assert [int(x) for x in d[13:16]] == ref[13:16]
    assert abs(float(d[16]) - ref[16]) <= 1e-9 and abs(float(d[17]) - ref[17]) <= 1e-7


def test_upgma_restatement_reproduces_example_tree(inputs):
    """the oracle's UPGMA restatement is pinned by the reference's own ExampleTree.nwk"""
    from oracle import oracle as O
    with open(inputs["g"]) as fh:
        parsed = M.Csv_to_dic_Roary(fh, ",", [], startcol=14)
    table = parsed["Roarydic"]
    tree = treemod.from_merges(table.strains, O.upgma_merges(table.matrix))
    assert treemod.to_scoary_newick(tree) == open(inputs["n"]).read().strip()


def test_bh_tie_rule():
    """methods.py:903-919: ties inherit from the less significant neighbour; last keeps its p."""
    p = np.array([0.001, 0.01, 0.01, 0.02, 0.5])
    m = 5
    bh = M.benjamini_hochberg(p, m)
    last = p[-1]
    want = [None] * 5
    want[4] = last
    want[3] = min(want[4], p[3] * m / 4.0)
    want[2] = min(want[3], p[2] * m / 3.0)
    want[1] = want[2]                         # tied with index 2
    want[0] = min(want[1], p[0] * m / 1.0)
    assert np.array_equal(bh, np.array(want))


def test_tree_utils_roundtrip_and_prune():
    t = [[["a", "b"], "c"], [["d", ["e", "f"]], "g"]]
    left, right, names = treemod.flatten(t)
    assert names == list("abcdefg") and len(left) == 6
    assert treemod.from_scoary_newick(treemod.to_scoary_newick(t)) == t
    assert treemod.prune(t, ["a", "b", None]) == ["c", [["d", ["e", "f"]], "g"]]
    assert treemod.prune(t, ["c"]) == [["a", "b"], [["d", ["e", "f"]], "g"]]
    assert treemod.prune(["a", "b"], ["a", "b"]) is None
    assert treemod.from_scoary_newick("((A:0.1,B:0.2):0.3,C:1);") == [["A", "B"], "C"]
    # support values / internal labels are discarded, polytomies resolved like ete3's
    # resolve_polytomy(recursive=True) (scoary/nwkhandler.py:19): [[..[[c(k-2), c(k-1)], c(k-3)].., c1], c0]
    assert treemod.from_newick("((a:0.1,b:0.2)0.95:0.3,(c:1,d:1)node7:0.5)root;") == [["a", "b"], ["c", "d"]]
    assert treemod.from_newick("(A,B,C);") == [["B", "C"], "A"]
    assert treemod.from_newick("(A:1,B:2,(C:1,D:1)100:3,E:0.5):0.0;") == [[[["C", "D"], "E"], "B"], "A"]
    assert treemod.from_newick("('it s'[&&NHX:x=1]:1,(\"q r\":2,(x)):3);") == ["it s", ["q r", "x"]]
    for bad in ["", "A;", "((A,B);", "(A,B));", "(A,'B);"]:
        with pytest.raises(ValueError):
            treemod.from_newick(bad)


def test_custom_tree_errors_exit_like_the_reference(tmp_path):
    """nwkhandler.py:16-17: sys.exit("Corrupted or non-existing custom tree file? ...")"""
    bad = tmp_path / "bad.nwk"
    bad.write_text("((A,B);")
    for path in (str(bad), str(tmp_path / "missing.nwk")):
        with pytest.raises(SystemExit) as ex:
            M.ReadTreeFromFile(path)
        assert "Corrupted or non-existing custom tree file?" in str(ex.value.code)
    ok = tmp_path / "ok.nwk"
    ok.write_text("((a:0.1,b:0.2)0.95:0.3,c:1,d:2);\n")
    assert M.ReadTreeFromFile(str(ok)) == ([["c", "d"], ["a", "b"]], ["c", "d", "a", "b"])


def test_errors_exit_like_the_reference(inputs, fake_engine):
    for argv, msg in [(["-g", inputs["g"]], "required"),
                      (["-g", inputs["g"], "-t", inputs["t"], "-e", "5"], "minimum number of permutations"),
                      (["-g", inputs["g"], "-t", inputs["t"], "-c", "P"], "without performing permutations"),
                      (["-g", inputs["g"], "-t", "/nonexistent.csv"], "Could not find the traits file"),
                      (["-g", inputs["g"], "-t", inputs["t"], "-p", "1.5"], "P must be between")]:
        with pytest.raises(SystemExit) as ex:
            M.main(argv=argv + ["-o", inputs["out"], "--no-time"])
        assert msg in str(ex.value.code)


def test_reference_style_dict_inputs(fake_engine):
    """Setup_results / ConvertUPGMAtoPhyloTree / Permute accept the reference's plain dicts."""
    strains = ["s%d" % i for i in range(12)]
    rng = np.random.default_rng(3)
    genedic = {}
    for g in range(6):
        row = {"Non-unique Gene name": "", "Annotation": "x"}
        row.update({s: int(rng.random() < 0.5) for s in strains})
        genedic["gene%d" % g] = row
    traitsdic = {"T": {s: str(int(rng.random() < 0.5)) for s in strains}}
    out = M.Setup_results(genedic, traitsdic, False)
    res = out["Results"]["T"]
    assert set(res) <= set(genedic) and all(k in next(iter(res.values())) for k in ("p_v", "B_p", "BH_p", "OR"))
    gtc = out["Gene_trait_combinations"]["T"][next(iter(res))]
    assert set(gtc.values()) <= {"AB", "Ab", "aB", "ab"} and list(gtc) == strains
    tree = [[["s0", "s1"], ["s2", "s3"]], [[["s4", "s5"], ["s6", "s7"]], [["s8", "s9"], ["s10", "s11"]]]]
    w = M.ConvertUPGMAtoPhyloTree(tree, gtc)
    assert set(w) == {"Total", "Pro", "Anti"} and w["Total"] >= max(w["Pro"], w["Anti"])
    emp = M.Permute(tree, gtc, 50, {"I": 0.05})
    assert 0.0 < emp <= 1.0


def _write_tricky_csv(path, G=120, N=70, seed=3):
    rng = np.random.default_rng(seed)
    hdr = M.ROARY_COLUMNS[:14] + ["Iso_%03d" % i for i in range(N)]
    with open(path, "w", newline="") as fh:
        fh.write(",".join('"%s"' % h for h in hdr) + "\r\n")
        for g in range(G):
            pres = rng.random(N) < rng.uniform(0.05, 0.95)
            cells = ['"locus_%d_%d"' % (g, j) if pres[j] else ['""', '0', '"-"', '', ' ""', '-'][(g + j) % 6] for j in range(N)]
            ann = '"hypothetical, protein ""%d"" x"' % g if g % 7 == 0 else '"annot %d"' % g
            fh.write('"gene_%d","nug%d",%s,"1","2","3","4","5","6","7","","8","9","10",' % (g, g, ann) + ",".join(cells)
                     + ("\r\n" if g % 2 else "\n"))


def test_native_csv_packer_equals_python_parser(tmp_path, monkeypatch, inputs):
    """sb_csv_pack_rows == the csv-module parser (which the golden CLI files pin),
    on the reference's example data and on a file with quoted commas, escaped quotes, mixed line ends,
    spaces after delimiters and every spelling of an absent cell."""
    tricky = str(tmp_path / "tricky.csv")
    _write_tricky_csv(tricky)
    for path, grab, allowed in [(inputs["g"], [], None), (tricky, [3, 10], None),
                                (tricky, [], {"Iso_%03d" % i: "all" for i in range(0, 70, 3)})]:
        monkeypatch.delenv("SCOARY_B200_PY_CSV", raising=False)
        with open(path) as fh:
            a = M.Csv_to_dic_Roary(fh, ",", list(grab), startcol=14, allowed_isolates=allowed)
        monkeypatch.setenv("SCOARY_B200_PY_CSV", "1")
        with open(path, newline="") as fh:
            b = M.Csv_to_dic_Roary(fh, ",", list(grab), startcol=14, allowed_isolates=allowed)
        ta, tb = a["Roarydic"], b["Roarydic"]
        assert ta.names == tb.names and ta.nugn == tb.nugn and ta.annotation == tb.annotation
        assert np.array_equal(ta.bits, tb.bits) and ta.extra == tb.extra and a["Strains"] == b["Strains"]
        assert np.array_equal(a["Zero_ones_matrix"], b["Zero_ones_matrix"])
    monkeypatch.delenv("SCOARY_B200_PY_CSV", raising=False)
    short = str(tmp_path / "short.csv")
    with open(tricky) as fh:
        lines = fh.read().splitlines()
    lines[5] = ",".join(lines[5].split(",")[:20])
    open(short, "w").write("\n".join(lines) + "\n")
    with pytest.raises(SystemExit) as ex, open(short) as fh:
        M.Csv_to_dic_Roary(fh, ",", [], startcol=14)
    assert "Could not read gene presence absence file" in str(ex.value.code)


def test_native_csv_packer_embedded_quotes_and_ragged_rows(tmp_path, monkeypatch):
    """a '"' inside an unquoted cell (`5" nuclease`) is a literal for csv.reader -- it must not swallow the
    row terminator in the native row scanner -- and rows wider than the header are left to the csv module."""
    import random
    rnd = random.Random(11)
    N = 9
    header = ",".join('"%s"' % c for c in M.ROARY_COLUMNS[:14]) + "," + ",".join("I%d" % j for j in range(N))
    tricky_ann = ['5" nuclease', 'x "quoted" y', '"starts quoted" tail"', '"a ""b"" c"', 'odd " count', '"multi\nline, cell"',
                  ' "after space"', 'plain', '3\'-5" exo"nuclease"', '""']
    for trial, wide_row in [(0, None), (1, 3), (2, None)]:
        lines = [header]
        if trial == 2:      # text that is not ASCII (byte offsets are not character offsets) and repeated identifiers
            tricky_ann = tricky_ann + ['na\u00efve "prot\u00e9ine" \u03b2', '"\u00e5 ""\u00f8"" \u00e6"']
        for g in range(40):
            ann = tricky_ann[g % len(tricky_ann)] if g % 2 == 0 else rnd.choice(tricky_ann)
            cells = [rnd.choice(['"l_%d"' % g, "l%d" % g, '', '0', '-', '""', 'a"b', '"-"']) for _ in range(N)]
            gid = g if trial < 2 or g % 9 else g // 18
            row = 'g%d,nug %d"x,%s,1,2,3,4,5,6,7,8,9,10,11,%s' % (gid, g, ann, ",".join(cells))
            if wide_row is not None and g == wide_row:
                row += ",extra"
            lines.append(row)
        path = str(tmp_path / ("quotes%d.csv" % trial))
        with open(path, "w", newline="", encoding="utf-8") as fh:
            fh.write("\r\n".join(lines[:20]) + "\n" + "\n".join(lines[20:]) + "\n")
        monkeypatch.delenv("SCOARY_B200_PY_CSV", raising=False)
        with open(path, encoding="utf-8") as fh:
            a = M.Csv_to_dic_Roary(fh, ",", [3], startcol=14)
        monkeypatch.setenv("SCOARY_B200_PY_CSV", "1")
        with open(path, newline="", encoding="utf-8") as fh:
            b = M.Csv_to_dic_Roary(fh, ",", [3], startcol=14)
        with open(path, newline="", encoding="utf-8") as fh:
            want = [r[0] for r in __import__("csv").reader(fh, skipinitialspace=True)][1:]
        ta, tb = a["Roarydic"], b["Roarydic"]
        assert len(want) == 40 and tb.names == [w for w in dict.fromkeys(want)]
        assert trial < 2 or (len(tb.names) < 40 and any(not x.isascii() for x in tb.annotation))
        assert ta.names == tb.names and ta.nugn == tb.nugn and ta.annotation == tb.annotation
        assert np.array_equal(ta.bits, tb.bits) and ta.extra == tb.extra
    monkeypatch.delenv("SCOARY_B200_PY_CSV", raising=False)


def test_native_csv_row_scan_in_parallel_pieces(tmp_path, monkeypatch):
    """Files above 4 MB are scanned in pieces cut at line feeds, each piece assuming that it starts a row; a line feed
    inside a quoted cell breaks the chain of pieces and the sequential scan decides.  Both ways: the csv module's table."""
    import csv as _csv
    N, G = 1000, 3200                                              # ~6.5 MB
    header = ",".join('"%s"' % c for c in M.ROARY_COLUMNS[:14]) + "," + ",".join("I%04d" % j for j in range(N))
    rng = np.random.default_rng(8)
    cells = np.where(rng.random((G, N)) < 0.4, "1", "0")
    for variant in ("plain", "embedded"):
        path = str(tmp_path / (variant + ".csv"))
        with open(path, "w", newline="") as fh:
            fh.write(header + "\n")
            for g in range(G):
                ann = "annot %d" % g
                if variant == "embedded" and g % 97 == 5:
                    ann = '"two\nlines, %d"' % g                   # every piece of the scan meets a few of these
                fh.write('g%d,,%s,1,2,3,4,5,6,7,8,9,10,11,' % (g, ann) + ",".join(cells[g]) + "\n")
        assert os.path.getsize(path) > (4 << 20)
        monkeypatch.delenv("SCOARY_B200_PY_CSV", raising=False)
        with open(path) as fh:
            a = M.Csv_to_dic_Roary(fh, ",", [], startcol=14)["Roarydic"]
        with open(path, newline="") as fh:
            rows = list(_csv.reader(fh, skipinitialspace=True))[1:]
        assert len(rows) == G and a.names == [r[0] for r in rows] and a.annotation == [r[2] for r in rows]
        want = np.asarray([[c == "1" for c in r[14:]] for r in rows], dtype=np.uint8)
        assert np.array_equal(a.matrix, want)
        # the two-pass form of the scan (count, then fill) and the one-pass form give the same row starts
        import ctypes
        from scoary_b200 import _lib
        lib = _lib.load()
        raw = open(path, "rb").read()
        n = lib.sb_csv_row_starts(raw, len(raw), b",", None, 0, None)
        two = np.empty(n, dtype=np.int64)
        assert lib.sb_csv_row_starts(raw, len(raw), b",", two.ctypes.data_as(ctypes.c_void_p), n, None) == n == G
        found = ctypes.c_void_p()
        assert lib.sb_csv_scan_rows(raw, len(raw), b",", ctypes.byref(found), None) == n
        one = np.ctypeslib.as_array(ctypes.cast(found, ctypes.POINTER(ctypes.c_int64)), shape=(n,)).copy()
        lib.sb_csv_free(found)
        assert np.array_equal(one, two) and raw[one[0]:one[0] + 3] == b"g0,"


@pytest.mark.parametrize("name,types", [("Example", None), ("generated", None), ("generated", "snp,del")])
def test_vcf2scoary_matches_reference_converter(name, types, tmp_path):
    """same output file as scoary/vcf2scoary.py (goldens written by the reference),
    and the direct VCF -> packed table path equals parsing that file."""
    from scoary_b200 import vcf2scoary as V
    gold = os.path.join(GOLD, "vcf")
    out = str(tmp_path / "out.csv")
    argv = ["--force", "--out", out] + (["--types", types] if types else []) + [os.path.join(gold, name + ".vcf")]
    with pytest.raises(SystemExit) as ex:
        V.main(argv)
    assert ex.value.code == 0
    want = os.path.join(gold, name + ("_snp_del" if types else "") + ".csv")
    assert open(out).read() == open(want).read()
    table = V.vcf_to_table(os.path.join(gold, name + ".vcf"), types.split(",") if types else "ALL")
    with open(want, newline="") as fh:
        ref = M.Csv_to_dic_Roary(fh, ",", [], startcol=10)["Roarydic"]
    # the reference's dict keeps one row per identifier (last wins): compare after the same de-duplication
    assert table.names == ref.names and table.strains == ref.strains[0:] and np.array_equal(table.bits, ref.bits)


def test_vcf_first_row_is_the_reference_ci_golden():
    """tests/test_scoary_output.py:16-17,123-136 of the reference"""
    import csv
    with open(os.path.join(GOLD, "vcf", "Example.csv")) as fh:
        rows = list(csv.reader(fh))
    assert rows[1] == ["NC_000962", "4013", "0", "T", "C", "9999", "0", "TYPE=snp", "GT", "False", "0", "1", "1", "1"]


def test_vectorised_binomial_test_is_bitwise_the_scalar_scipy_call():
    """PairWiseComparisons evaluates ss.binom_test(k, n, 0.5) (scoary/methods.py:1267-1275) for all genes in one
    vectorised SciPy call; the columns must not change by a bit."""
    from scipy import stats as ss
    rng = np.random.default_rng(3)
    pairs = [(k, n) for n in range(1, 40) for k in range(n + 1)]
    for n in rng.integers(40, 5001, 60).tolist():
        pairs += [(int(k), n) for k in set(rng.integers(0, n + 1, 4).tolist() + [0, n, n // 2, (n + 1) // 2])]
    k = np.array([p[0] for p in pairs] + [0]), np.array([p[1] for p in pairs] + [0])
    M._BINOM_CACHE.clear()
    got = M._binom_two_sided_many(*k)
    want = np.array([float(ss.binomtest(a, b, 0.5).pvalue) for a, b in pairs])
    assert np.array_equal(got[:-1].view(np.uint64), want.view(np.uint64))
    assert np.isnan(got[-1])                 # (0, 0): nan, as the reference prints with this SciPy
    assert np.array_equal(M._binom_two_sided_many(*k)[:-1].view(np.uint64), want.view(np.uint64))      # from the cache


def test_early_stop_table_is_the_reference_expression():
    """rmin[i] found around the boundary must equal the first r of the reference's own test
    1 - ss.binom.cdf(r, i, 0.1) < 0.05 (scoary/methods.py:1360-1361) scanned from r = 0."""
    from scipy import stats as ss
    P = 700
    M._RMIN_CACHE.clear()
    got = M.early_stop_table(P)
    for i in range(30, P):
        r = np.arange(0, i + 2)
        assert got[i] == int(np.argmax((1 - ss.binom.cdf(r, i, 0.1)) < 0.05)), i
    assert np.all(got[:30] == np.iinfo(np.int32).max)
    assert [int(M.early_stop_table(10000)[i]) for i in (30, 50, 100, 1000, 9999)] == [6, 9, 15, 116, 1049]


@pytest.mark.parametrize("vname,tname,traits", [("generated", "generated_traits.csv", ["resistant", "with_missing"]),
                                                ("Example", "ExampleVCFTrait.csv", ["ExampleVCFtrait"])])
def test_cli_reads_a_vcf_directly(vname, tname, traits, tmp_path, fake_engine):
    """`-g x.vcf` = the reference's vcf2scoary followed by `scoary -g converted.csv -s 11` (goldens written
    by the reference that way), without the CSV round trip."""
    vdir = os.path.join(GOLD, "vcf")
    out = str(tmp_path / "out")
    _run(["-g", os.path.join(vdir, vname + ".vcf"), "-t", os.path.join(vdir, tname), "-p", "1.0", "-c", "I", "-o", out,
          "--no-time"])
    for t in traits:
        assert _read(os.path.join(out, t + ".results.csv")) == \
            _read(os.path.join(vdir, "cli_" + vname, t + ".results.csv")), t


def test_native_vcf_packer_equals_python_parser(tmp_path, monkeypatch):
    """csrc/vcf_pack.cpp against the csv-module parser on generated files: line endings, empty lines,
    multi-allelic sites with multi-digit and zero-padded alleles, '.', TYPE= filters, -r subsets, and
    lines only the Python parser may judge (then the native path must step aside)."""
    import random
    from scoary_b200 import vcf2scoary as V
    rng = random.Random(7)
    used = 0
    for it in range(40):
        ns = rng.choice([0, 1, 5, 63, 64, 65, 130])
        samples = ["S%03d" % i for i in range(ns)]
        lines = ["##fileformat=VCFv4.2", '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">', "",
                 "\t".join(["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + samples)]
        for k in range(rng.choice([0, 1, 7, 60])):
            nalt = rng.choice([1, 1, 1, 2, 3, 12])
            info = rng.choice(["TYPE=snp", "DP=3;TYPE=ins", "SUBTYPE=del;X=1", "TYPE=;TYPE=complex", "NOTYPE", "TYPE=snp_x;Q"])
            cells = [rng.choice(["0", "0", "1", ".", "2", "3", "12", "01", "-", "", "10"]) +
                     rng.choice(["", ":35", ":1:2", ":"]) for _ in samples]
            if it % 8 == 7 and k == 3 and cells:
                cells[0] = "1/1"
            if rng.random() < 0.05:
                lines.append("")
            lines.append("\t".join(["chr%d" % (k % 3), str(100 + k * 7), rng.choice(["id%d" % k, ".", ""]), "A",
                                    ",".join("ACGT"[a % 4] * (1 + a // 4) for a in range(nalt)), "99", "PASS", info,
                                    "GT:DP"] + cells))
        path = str(tmp_path / ("f%d.vcf" % it))
        eol = rng.choice(["\n", "\r\n", "\r"])
        with open(path, "w", newline="") as fh:
            fh.write(eol.join(lines) + (eol if it % 3 else ""))
        types = rng.choice(["ALL", ["snp"], ["snp", "ins", "complex"], ["del"], ["snp_x"]])
        keep = rng.choice([None, set(samples[::2])])

        def run(py):
            if py:
                monkeypatch.setenv("SCOARY_B200_PY_CSV", "1")
            else:
                monkeypatch.delenv("SCOARY_B200_PY_CSV", raising=False)
            try:
                return V.vcf_to_table(path, types, keep)
            except SystemExit as ex:
                return "exit %r" % (ex.code,)
        monkeypatch.delenv("SCOARY_B200_PY_CSV", raising=False)
        try:
            used += V._native_table(path, types, keep) is not None
        except SystemExit:
            pass
        a, b = run(False), run(True)
        if isinstance(a, str) or isinstance(b, str):
            assert a == b, (it, a, b)
            continue
        assert (a.names, a.nugn, a.annotation, a.strains) == (b.names, b.nugn, b.annotation, b.strains), it
        assert a.bits.shape == b.bits.shape and np.array_equal(a.bits, b.bits), it
    assert used >= 20


def test_duplicate_identifiers_count_in_the_tree_but_not_in_the_table(tmp_path, fake_engine):
    """A converted VCF table repeats CHROM_|_POS_|_ID for every allele of a multi-allelic site: the reference's
    gene dictionary keeps the last such line (methods.py:458-463) while its distance matrix is built from every
    line (:496).  Golden = the reference CLI on the same CSV."""
    vdir = os.path.join(GOLD, "vcf")
    out = str(tmp_path / "out")
    _run(["-g", os.path.join(vdir, "generated.csv"), "-s", "11", "-t", os.path.join(vdir, "generated_traits.csv"),
          "-p", "1.0", "-c", "I", "-o", out, "--no-time"])
    for t in ("resistant", "with_missing"):
        assert _read(os.path.join(out, t + ".results.csv")) == \
            _read(os.path.join(vdir, "cli_generated", t + ".results.csv")), t


MORE_SCENARIOS = {
    "bh_pw": ["-c", "I", "BH", "PW", "-p", "0.05", "0.2", "0.05", "-m", "40"],
    "grabcols": ["--include_input_columns", "4,6-8", "-p", "0.01"],
}


@pytest.mark.parametrize("name", list(MORE_SCENARIOS))
def test_cli_more_reference_scenarios(name, inputs, fake_engine):
    """BH + pairwise cut-offs with -m, and --include_input_columns: text-identical to the reference."""
    _run(["-g", inputs["g"], "-t", inputs["t"], "-o", inputs["out"], "--no-time"] + MORE_SCENARIOS[name])
    for trait in ("Tetracycline_resistance", "Bogus_trait"):
        assert _read(os.path.join(inputs["out"], trait + ".results.csv")) == \
            _read(os.path.join(GOLD, name, trait + ".results.csv")), (name, trait)


def test_cli_writes_the_reduced_gene_table(inputs, fake_engine):
    """-r with -w: the reduced presence/absence file and the results computed from it (methods.py:510-544)."""
    _run(["-g", inputs["g"], "-t", inputs["t"], "-o", inputs["out"], "--no-time", "-r", inputs["r"], "-w", "-p", "0.01",
          "--no_pairwise"])
    assert _read(os.path.join(inputs["out"], "gene_presence_absence_reduced.csv")) == \
        _read(os.path.join(GOLD, "reduced", "gene_presence_absence_reduced.csv.gz"))
    for trait in ("Tetracycline_resistance", "Bogus_trait"):
        assert _read(os.path.join(inputs["out"], trait + ".results.csv")) == \
            _read(os.path.join(GOLD, "reduced", trait + ".results.csv"))


def test_cli_semicolon_delimiter(tmp_path, fake_engine):
    """--delimiter ';' applies to both inputs and to the result files (methods.py:350-351, :1159-1197)."""
    sdir = os.path.join(GOLD, "semicolon")
    g = tmp_path / "genes_semicolon.csv"
    with gzip.open(os.path.join(sdir, "genes_semicolon.csv.gz"), "rb") as fi, open(g, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    out = str(tmp_path / "out")
    _run(["-g", str(g), "-t", os.path.join(sdir, "traits_semicolon.csv"), "--delimiter", ";", "-p", "0.05", "-o", out,
          "--no-time"])
    for trait in ("Tetracycline_resistance", "Bogus_trait"):
        assert _read(os.path.join(out, trait + ".results.csv")) == _read(os.path.join(sdir, trait + ".results.csv"))


def test_prune_removes_the_none_clusters_of_a_degenerate_upgma():
    """With tied distances the reference's upgma can return a tree with None in place of a cluster
    (methods.py:685-686); Prunedic always ends with None (:612) and PruneForMissing (:709-739) removes it."""
    broken = [[[["a", "b"], "c"], "d"], None]
    assert treemod.prune(broken, [None]) == [[["a", "b"], "c"], "d"]
    assert treemod.prune(broken, ["d", None]) == [["a", "b"], "c"]
    assert treemod.prune([["a", None], [None, None]], [None]) == "a"
    assert treemod.to_scoary_newick(broken) == "(((('a', 'b'), 'c'), 'd'), None);"


def test_cli_edge_cases_found_by_the_fuzzer(tmp_path, fake_engine):
    """tests/golden/edge: one variable gene (the reference's upgma leaves a None cluster that PruneForMissing
    removes; Tree.nwk still shows it), an all-zero trait (odds ratio nan), no contrasting pairs (binomial p nan),
    missing values.  Text-identical to the reference."""
    edir = os.path.join(GOLD, "edge")
    out = str(tmp_path / "out")
    _run(["-g", os.path.join(edir, "genes.csv"), "-t", os.path.join(edir, "traits.csv"), "-p", "1.0", "-c", "I", "-u",
          "-o", out, "--no-time"])
    for f in ("perfect", "allzero", "withNA", "mixed"):
        assert _read(os.path.join(out, f + ".results.csv")) == _read(os.path.join(edir, f + ".results.csv")), f
    assert _read(os.path.join(out, "Tree.nwk")) == _read(os.path.join(edir, "Tree.nwk"))


def test_lazy_zero_ones_matrix_matches_the_unpacked_cells():
    """Zero_ones_matrix of large tables / VCF input (reference: methods.py:496-497) is unpacked on first use only."""
    from scoary_b200 import methods as M
    rng = np.random.default_rng(5)
    cells = (rng.random((40, 70)) < 0.4).astype(np.uint8)                 # genes x isolates
    table = M.GeneTable([f"g{i}" for i in range(40)], [""] * 40, [""] * 40, [f"s{j}" for j in range(70)], matrix=cells)
    rows = [3, 7, 8, 30]
    z = M.LazyZeroOnes(table, rows)
    assert z.shape == (70, 4) and len(z) == 70 and z._m is None
    assert np.array_equal(np.asarray(z), cells[rows].T)
    assert np.array_equal(z[5], cells[rows, 5]) and np.array_equal(np.stack(list(z)), cells[rows].T)
