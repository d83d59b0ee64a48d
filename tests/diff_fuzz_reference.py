"""TEST INFRASTRUCTURE (build container only: needs /root/reference).

Differential fuzzer: random small inputs and command lines go through the UNMODIFIED reference CLI
(oracle/ref_shim.py) and through this repo's CLI with the oracle-backed FakeEngine; every file either
one writes must be byte-identical, and both must exit the same way.  tests/test_diff_fuzz.py runs a
few seeds of it; `python tests/diff_fuzz_reference.py 0 200` runs a campaign.

Not covered (the reference is not deterministic or not runnable there): -e (unseeded RNG), -n (ete3)."""
import contextlib
import io
import os
import random
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import ref_shim  # noqa: E402

ROARY = ["Gene", "Non-unique Gene name", "Annotation", "No. isolates", "No. sequences", "Avg sequences per isolate",
         "Genome Fragment", "Order within Fragment", "Accessory Fragment", "Accessory Order with Fragment", "QC",
         "Min group size nuc", "Max group size nuc", "Avg group size nuc"]


def make_case(rng, d):
    n = rng.choice([4, 5, 8, 13, 24, 40])
    g = rng.choice([1, 3, 12, 60, 150])
    delim = rng.choice([",", ",", ",", ";"])
    iso = ["iso%02d" % j if rng.random() < 0.8 else "S.%d-x" % j for j in range(n)]
    roary = rng.random() < 0.7
    head = list(ROARY) if roary else ["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT", "DUMMY"]
    start = len(head) + 1

    def q(x):
        return '"%s"' % x

    clade = [rng.random() < 0.5 for _ in iso]
    lines = [delim.join(q(h) for h in head + iso)]
    names = []
    for k in range(g):
        f = rng.choice([0.0, 0.05, 0.3, 0.5, 0.8, 1.0]) if rng.random() < 0.5 else rng.random()
        mode = rng.random()
        bits = [(c if mode < 0.3 else (rng.random() < f)) for c in clade]
        if mode < 0.3 and rng.random() < 0.5:
            bits = [b ^ (rng.random() < 0.1) for b in bits]
        if k and rng.random() < 0.15:
            bits = prev                                    # identical pattern (collapse)
        elif k and rng.random() < 0.05:
            bits = [not x for x in prev]                   # complementary pattern
        prev = bits
        name = "gene%d" % k if not (names and rng.random() < 0.05) else rng.choice(names)    # repeated identifier
        names.append(name)
        cells = [rng.choice(["x_%d" % k, "1", "a b", "0.0", "--", "00", 'q""q', "x%sy" % delim]) if b
                 else rng.choice(["", "", "0", "-"]) for b in bits]
        if roary:
            lead = [name, rng.choice(["", "nug%d" % k]), rng.choice(["hypothetical protein", "a, b", "", 'said ""hi""'])] + \
                   [str(rng.randint(1, 9)) for _ in head[3:]]
        else:
            lead = ["chr%d" % (k % 2), str(10 * k), rng.choice([".", "v%d" % k])] + ["A", "C", "9", "PASS", "TYPE=snp", "GT",
                                                                                   rng.choice(["True", "False"])]
        sep = delim + (" " if rng.random() < 0.1 else "")          # skipinitialspace
        lines.append(sep.join((q(c) if (rng.random() < 0.9 or delim in c or '"' in c) else c) for c in lead + cells))
    gpath = os.path.join(d, "genes.csv")
    with open(gpath, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    nt = rng.choice([1, 1, 2, 3])
    tl = [delim.join([rng.choice(["", "", "Name"])] + ["trait%d" % t for t in range(nt)])]
    order = list(range(n))
    if rng.random() < 0.5:
        rng.shuffle(order)
    if rng.random() < 0.15 and n > 5:
        order = order[:-1]                                # an isolate missing from the traits file
    for j in order:
        vals = []
        for t in range(nt):
            r = rng.random()
            vals.append(rng.choice(["NA", "-", "."]) if r < 0.08 else ("1" if (clade[j] ^ (rng.random() < 0.25)) else "0"))
        tl.append(delim.join([iso[j]] + vals))
    tpath = os.path.join(d, "traits.csv")
    with open(tpath, "w") as fh:
        fh.write("\n".join(tl) + "\n")
    argv = ["-g", gpath, "-t", tpath, "-s", str(start)]
    if delim != ",":
        argv += ["--delimiter", delim]
    corr = rng.choice([["I"], ["I"], ["B"], ["BH"], ["I", "EPW"], ["I", "PW"], ["BH", "PW", "EPW"], ["I", "B", "BH"]])
    nopw = rng.random() < 0.2
    if nopw:
        corr = [c for c in corr if c not in ("PW", "EPW")] or ["I"]
        argv.append("--no_pairwise")
    pv = [rng.choice(["1.0", "0.5", "0.05", "0.9"])] if rng.random() < 0.5 else [rng.choice(["1.0", "0.6", "0.2"]) for _ in corr]
    argv += ["-c"] + corr + ["-p"] + pv
    if rng.random() < 0.3:
        argv.append("--collapse")
    if rng.random() < 0.3 and not nopw:
        # permutations: the reference's RNG is unseeded, so both sides get the same stand-in for the label
        # shuffling (see fake_hits) and everything around it -- columns, sorting, the P cut-off -- is compared
        argv += ["-e", str(rng.choice([10, 50, 400]))]
        if rng.random() < 0.5:
            i = argv.index("-c")
            argv.insert(i + 1 + len(corr), "P")
            j = argv.index("-p")
            if len(pv) > 1:
                argv.insert(j + 1 + len(pv), rng.choice(["1.0", "0.5", "0.11"]))
    if rng.random() < 0.3:
        argv += ["-m", str(rng.choice([1, 3, 10, 1000]))]
    if rng.random() < 0.3 and not nopw:
        argv.append("-u")
    if rng.random() < 0.25 and n >= 8:
        rpath = os.path.join(d, "restrict.csv")
        keep = [s for s in iso if rng.random() < 0.7] or iso[:4]
        with open(rpath, "w") as fh:
            fh.write(",".join(keep) + "\n")
        argv += ["-r", rpath]
        if rng.random() < 0.3:
            argv.append("-w")
    if rng.random() < 0.2 and roary:
        argv += ["--include_input_columns", rng.choice(["4", "4,6-8", "ALL", "5-7,10"])]
    return argv


def make_vcf_case(rng, d):
    """A random VCF; the reference converts it (vcf2scoary) and reads the CSV with -s 11, ours reads the VCF."""
    n = rng.choice([4, 6, 9, 20, 66])
    samples = ["S%02d" % j for j in range(n)]
    clade = [rng.random() < 0.5 for _ in samples]
    vl = ["##fileformat=VCFv4.2", '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
          '##INFO=<ID=TYPE,Number=A,Type=String,Description="The type of allele.">',
          "\t".join(["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + samples)]
    for k in range(rng.choice([1, 5, 40, 120])):
        nalt = rng.choice([1, 1, 1, 2, 3])
        cells = []
        for j in range(n):
            if rng.random() < 0.4:
                g = str(rng.randint(1, nalt)) if (clade[j] ^ (rng.random() < 0.15)) else "0"
            else:
                g = rng.choice(["0", "0", "1", ".", str(rng.randint(0, nalt))])
            cells.append(g + rng.choice(["", ":%d" % rng.randint(1, 99)]))
        vl.append("\t".join(["chr%d" % (k % 2), str(100 + 3 * k), rng.choice([".", "id%d" % k]), "A",
                             ",".join(rng.sample("CGT", nalt)), "50", "PASS", "TYPE=" + rng.choice(["snp", "ins", "del"]),
                             "GT:DP"] + cells))
    vpath = os.path.join(d, "variants.vcf")
    with open(vpath, "w") as fh:
        fh.write("\n".join(vl) + "\n")
    tpath = os.path.join(d, "traits.csv")
    with open(tpath, "w") as fh:
        fh.write(",t0,t1\n")
        for j, smp in enumerate(samples):
            fh.write("%s,%d,%s\n" % (smp, clade[j] ^ (rng.random() < 0.2), rng.choice(["0", "1", "1", "NA"])))
    common = ["-t", tpath, "-c", "I", "-p", rng.choice(["1.0", "0.3"])]
    if rng.random() < 0.3:
        common.append("--collapse")
    if rng.random() < 0.3:
        common.append("--no_pairwise")
    return vpath, common


def convert_with_reference(vpath, cpath):
    import importlib
    ref_shim.load()
    conv = importlib.import_module("scoary.vcf2scoary")
    old = sys.argv
    sys.argv = ["vcf2scoary", "--force", "--out", cpath, vpath]
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            conv.main()
    except SystemExit:
        pass
    finally:
        sys.argv = old


def fake_hits(total, pro, anti, P):
    """Deterministic stand-in for the number of permutations that beat the observed statistic (a function of
    the unpermuted walk, which both sides compute on the same tree)."""
    return (total * 7 + pro * 13 + anti * 31 + 3) % (P + 1)


def _ref_permute(tree, GTC, permutations, cutoffs):
    w = ref_shim.load().ConvertUPGMAtoPhyloTree(tree, GTC)
    return (fake_hits(w["Total"], w["Pro"], w["Anti"], permutations) + 1.0) / (permutations + 1.0)


def run_reference(argv, out):
    ref_shim.load().Permute = _ref_permute
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        try:
            return ref_shim.run_cli(argv + ["-o", out, "--no-time"])
        except Exception as e:
            return "crash: %s: %s" % (type(e).__name__, e)


def run_ours(argv, out):
    import numpy as np
    from fake_engine import FakeEngine
    from scoary_b200 import methods as M

    class Engine(FakeEngine):
        def permute(self, t, P, seed=0, gene_idx=None, early_stop=False, rmin=None):
            pairs = self.pairwise(t, gene_idx)
            r = [fake_hits(int(a), int(b), int(c), P) for a, b, c in pairs]
            return pairs, np.asarray(r, dtype=np.int32), np.full(len(r), P, dtype=np.int32)

    M._ENGINE = Engine()
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        try:
            M.main(argv=argv + ["-o", out, "--no-time"])
        except SystemExit as e:
            return e.code
        except Exception as e:                      # a crash of ours is a finding, not a fuzzer failure
            return "crash: %s: %s" % (type(e).__name__, e)
    return 0


def files(out):
    res = {}
    if os.path.isdir(out):
        for f in sorted(os.listdir(out)):
            if not f.endswith(".log"):
                with open(os.path.join(out, f)) as fh:
                    res[f] = fh.read()
    return res


def one(seed, keep=False):
    rng = random.Random(seed)
    d = tempfile.mkdtemp(prefix="scoary_fuzz_%d_" % seed)
    try:
        if rng.random() < 0.2:
            vpath, common = make_vcf_case(rng, d)
            cpath = os.path.join(d, "converted.csv")
            convert_with_reference(vpath, cpath)
            argv = ["-g", vpath, "-s", "11"] + common
            a = run_reference(["-g", cpath, "-s", "11"] + common, os.path.join(d, "ref"))
            b = run_ours(["-g", vpath] + common, os.path.join(d, "ours"))
        else:
            argv = make_case(rng, d)
            a = run_reference(argv, os.path.join(d, "ref"))
            b = run_ours(argv, os.path.join(d, "ours"))
        fa, fb = files(os.path.join(d, "ref")), files(os.path.join(d, "ours"))
        ok_exit = (a in (0, None)) == (b in (0, None))
        problems = []
        if isinstance(a, str) and a.startswith("crash"):
            # the reference itself dies with a traceback (e.g. IndexError at methods.py:914 when a trait has no
            # testable gene): nothing to compare, and not a behaviour to reproduce
            problems = []
        elif not ok_exit:
            problems.append("exit: reference %r, ours %r" % (a, b))
        elif a in (0, None):
            if sorted(fa) != sorted(fb):
                problems.append("files: %s vs %s" % (sorted(fa), sorted(fb)))
            for f in fa:
                if f in fb and fa[f] != fb[f]:
                    problems.append("content of " + f)
        return argv, problems, d
    finally:
        if not keep:
            shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    lo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    hi = int(sys.argv[2]) if len(sys.argv) > 2 else lo + 50
    bad = 0
    for seed in range(lo, hi):
        argv, problems, d = one(seed)
        if problems:
            bad += 1
            print("seed %d: %s\n    %s" % (seed, "; ".join(problems), " ".join(argv[4:])))
    print("%d seeds, %d differ" % (hi - lo, bad))
