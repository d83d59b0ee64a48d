#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED
reference (/root/reference, imported through oracle/ref_shim.py) in the build
container.  The reference cannot travel to the GPU box, so its outputs are
committed here together with this script.

    python tests/golden/make_golden.py

Fixtures written:
  inputs/                      the reference's example data (input DATA, gzipped)
  <scenario>/<Trait>.results.csv[.gz]   full results of the reference CLI
       default     .travis.yml:18  (-c I -p 0.05)
       nopairwise  .travis.yml:20  (--no_pairwise)
       restrict    .travis.yml:22  (-r Restrict_to.csv), all genes (-p 1.0)
       advanced    .travis.yml:24  (-p 0.01 1E-5 -c B EPW --collapse -m 50 -u; tree built internally)
       all         every gene, both traits (-p 1.0 -c I)  -> pins counts, p, B_p, BH_p, pairs, binomial p
       collapse    every gene with --collapse (-p 1.0)
       perm        -e 200 -c I EPW -p 0.05 0.05 with random.seed(1): Empirical_p (statistical pin only)
  walks.json       random trees x gene-trait combinations -> reference ConvertUPGMAtoPhyloTree
  upgma.json       random presence matrices -> reference CreateTriangularDistanceMatrix/QuadTree/upgma tree
  fisher.json      2x2 tables -> scipy.stats.fisher_exact as the reference calls it (methods.py:854)
  vcf/*.csv        reference vcf2scoary output for Example.vcf and a generated multi-allelic VCF
  tetrcg_first_row.json   the reference's own CI golden (tests/test_scoary_output.py:12-14)
"""
import gzip
import json
import os
import random
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

EX = os.path.join(ref_shim.REFERENCE_ROOT, "scoary", "exampledata")
G = os.path.join(EX, "Gene_presence_absence.csv")
T = os.path.join(EX, "Tetracycline_resistance.csv")
R = os.path.join(EX, "Restrict_to.csv")

SCENARIOS = {
    "default": (["-g", G, "-t", T], False),
    "nopairwise": (["-g", G, "-t", T, "--no_pairwise"], False),
    "restrict": (["-g", G, "-t", T, "-r", R, "-p", "1.0"], True),
    "advanced": (["-g", G, "-t", T, "-p", "0.01", "1E-5", "-c", "B", "EPW", "--collapse", "-m", "50", "-u"], False),
    "all": (["-g", G, "-t", T, "-p", "1.0", "-c", "I"], True),
    "collapse": (["-g", G, "-t", T, "-p", "1.0", "-c", "I", "--collapse"], True),
    "perm": (["-g", G, "-t", T, "-e", "200", "-c", "I", "EPW", "-p", "0.05", "0.05"], False),
    # behaviours the Travis scenarios do not reach
    # (-n needs ete3, which this container does not have: custom trees are checked against the UPGMA run instead)
    "bh_pw": (["-g", G, "-t", T, "-c", "I", "BH", "PW", "-p", "0.05", "0.2", "0.05", "-m", "40"], False),
    "grabcols": (["-g", G, "-t", T, "--include_input_columns", "4,6-8", "-p", "0.01"], False),
    "reduced": (["-g", G, "-t", T, "-r", R, "-w", "-p", "0.01", "--no_pairwise"], False),
}


def store(src, dst, gz):
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    if gz:
        with open(src, "rb") as fi, gzip.GzipFile(dst + ".gz", "wb", mtime=0) as fo:
            shutil.copyfileobj(fi, fo)
    else:
        shutil.copyfile(src, dst)


def main():
    m = ref_shim.load()
    # ---- input data (not code): needed by the GPU-box CLI parity tests
    for f in ("Gene_presence_absence.csv", "Tetracycline_resistance.csv", "Restrict_to.csv", "ExampleTree.nwk"):
        store(os.path.join(EX, f), os.path.join(HERE, "inputs", f), gz=f.startswith("Gene"))
    # ---- CLI scenarios
    for name, (argv, gz) in SCENARIOS.items():
        out = tempfile.mkdtemp(prefix="golden_%s_" % name)
        random.seed(1)
        rc = ref_shim.run_cli(argv + ["-o", out, "--no-time"])
        assert rc in (0, None), (name, rc)
        for f in sorted(os.listdir(out)):
            if f.endswith(".results.csv") or f.endswith(".nwk") or f.startswith("gene_presence_absence_reduced"):
                store(os.path.join(out, f), os.path.join(HERE, name, f),
                      (gz and f.endswith(".csv")) or f.startswith("gene_presence_absence_reduced"))
        shutil.rmtree(out)
        print("scenario", name, "done")
    # ---- edge cases found by tests/diff_fuzz_reference.py, pinned as a fixed scenario:
    #      one variable gene (tied distances: the reference's upgma leaves a None cluster, PruneForMissing removes it),
    #      a trait that is all zeros (SciPy: odds ratio nan, p 1), a gene/trait pair without contrasting pairs
    #      (binomial p nan), missing values, an identifier used twice
    edir = os.path.join(HERE, "edge")
    os.makedirs(edir, exist_ok=True)
    iso = ["iso%02d" % j for j in range(8)]
    roary = ["Gene", "Non-unique Gene name", "Annotation", "No. isolates", "No. sequences", "Avg sequences per isolate",
             "Genome Fragment", "Order within Fragment", "Accessory Fragment", "Accessory Order with Fragment", "QC",
             "Min group size nuc", "Max group size nuc", "Avg group size nuc"]
    rows = [("geneA", "00010111"), ("core", "11111111"), ("absent", "00000000")]
    with open(os.path.join(edir, "genes.csv"), "w") as fh:
        fh.write(",".join('"%s"' % c for c in roary + iso) + "\n")
        for name, bitsx in rows:
            fh.write(",".join('"%s"' % c for c in [name, "", "edge case"] + ["1"] * 11 +
                              [("x" if ch == "1" else "") for ch in bitsx]) + "\n")
    with open(os.path.join(edir, "traits.csv"), "w") as fh:
        fh.write(",perfect,allzero,withNA,mixed\n")
        for j, nm in enumerate(iso):
            fh.write("%s,%s,0,%s,%s\n" % (nm, "00010111"[j], "NA" if j in (1, 6) else "01010011"[j], "10010110"[j]))
    out = tempfile.mkdtemp(prefix="golden_edge_")
    rc = ref_shim.run_cli(["-g", os.path.join(edir, "genes.csv"), "-t", os.path.join(edir, "traits.csv"), "-p", "1.0",
                           "-c", "I", "-u", "-o", out, "--no-time"])
    assert rc in (0, None), rc
    for f in sorted(os.listdir(out)):
        if f.endswith(".results.csv") or f.endswith(".nwk"):
            store(os.path.join(out, f), os.path.join(edir, f), False)
    shutil.rmtree(out)
    # ---- --delimiter ';' : the example files rewritten with semicolons (inputs stored gzipped), results from the reference
    import csv
    sdir = os.path.join(HERE, "semicolon")
    os.makedirs(sdir, exist_ok=True)
    tmpd = tempfile.mkdtemp(prefix="golden_semi_")
    for src, dst in ((G, "genes_semicolon.csv"), (T, "traits_semicolon.csv")):
        with open(src, newline="") as fi, open(os.path.join(tmpd, dst), "w", newline="") as fo:
            w = csv.writer(fo, delimiter=";", lineterminator="\n")
            for k, row in enumerate(csv.reader(fi, skipinitialspace=True)):
                w.writerow(row if (k < 1200 or dst.startswith("traits")) else row)
        store(os.path.join(tmpd, dst), os.path.join(sdir, dst), gz=dst.startswith("genes"))
    out = tempfile.mkdtemp(prefix="golden_semi_out_")
    rc = ref_shim.run_cli(["-g", os.path.join(tmpd, "genes_semicolon.csv"), "-t", os.path.join(tmpd, "traits_semicolon.csv"),
                           "--delimiter", ";", "-p", "0.05", "-o", out, "--no-time"])
    assert rc in (0, None), rc
    for f in sorted(os.listdir(out)):
        if f.endswith(".results.csv"):
            store(os.path.join(out, f), os.path.join(sdir, f), False)
    shutil.rmtree(out)
    shutil.rmtree(tmpd)
    # ---- direct calls: PhyloTree walks
    rng = random.Random(20260924)

    def rand_tree(names):
        nodes = list(names)
        while len(nodes) > 1:
            a = nodes.pop(rng.randrange(len(nodes)))
            b = nodes.pop(rng.randrange(len(nodes)))
            nodes.append([a, b])
        return nodes[0]

    def comb_tree(names):
        t = names[0]
        for n in names[1:]:
            t = [t, n]
        return t

    walks = []
    for it in range(400):
        n = rng.choice([2, 3, 4, 5, 7, 8, 16, 33, 64, 100, 150])
        names = ["i%d" % k for k in range(n)]
        tree = comb_tree(names) if it % 10 == 0 else rand_tree(names)
        pg, pt = rng.random(), rng.random()
        gtc = {nm: ("A" if rng.random() < pg else "a") + ("B" if rng.random() < pt else "b") for nm in names}
        ref = m.ConvertUPGMAtoPhyloTree(tree, gtc)
        walks.append({"tree": tree, "gtc": gtc, "out": [ref["Total"], ref["Pro"], ref["Anti"]]})
    with open(os.path.join(HERE, "walks.json"), "w") as fh:
        json.dump(walks, fh, separators=(",", ":"))
    # ---- Fisher tables as the reference calls SciPy
    import numpy as np
    from scipy import stats as ss
    npr = np.random.default_rng(20260924)
    tabs = []
    for it in range(600):
        N = int(npr.choice([7, 20, 100, 1000, 5000, 10000]))
        r1, c1 = int(npr.integers(1, N)), int(npr.integers(1, N))
        if it % 3 == 0:
            r1 = N // 2
        if it % 6 == 0:
            c1 = N // 2
        lo, hi = max(0, r1 + c1 - N), min(r1, c1)
        a = int(npr.integers(lo, hi + 1))
        if it % 2 == 0:
            sd = max(r1 * c1 / N * (1 - r1 / N) * (1 - c1 / N), 1.0) ** 0.5
            a = int(np.clip(round(r1 * c1 / N + npr.normal() * 3 * sd), lo, hi))
        tpgp, tpgn, tngp, tngn = a, r1 - a, c1 - a, N - r1 - c1 + a
        odds, p = ss.fisher_exact([[tpgp, tpgn], [tngp, tngn]])
        tabs.append({"tpgp": tpgp, "tngp": tngp, "tpgn": tpgn, "tngn": tngn, "p": float(p).hex(),
                     "odds": None if not np.isfinite(odds) else float(odds).hex()})
    with open(os.path.join(HERE, "fisher.json"), "w") as fh:
        json.dump({"scipy": __import__("scipy").__version__, "tables": tabs}, fh, separators=(",", ":"))
    # ---- tree construction: reference upgma on random matrices (many tied distances)
    upg = []
    nrng = np.random.default_rng(77)
    for it in range(60):
        n = int(nrng.choice([2, 3, 4, 5, 8, 13, 16, 17, 31, 40]))
        g = int(nrng.choice([3, 8, 30, 200]))
        mat = (nrng.random((g, n)) < nrng.uniform(0.1, 0.9)).astype(int)
        if it % 4 == 0 and n > 3:
            mat[:, 1] = mat[:, 0]
            mat[:, n - 1] = mat[:, 0]
        var = [row for row in mat.tolist() if 0 < sum(row) < n]
        if not var:
            continue
        names = ["s%d" % k for k in range(n)]
        zom = list(map(list, zip(*var)))                      # isolates x variable genes (methods.py:502)
        tdm = m.CreateTriangularDistanceMatrix(zom, names)
        tree = m.upgma(m.PopulateQuadTreeWithDistances(tdm))
        upg.append({"matrix": mat.tolist(), "tree": tree})
    with open(os.path.join(HERE, "upgma.json"), "w") as fh:
        json.dump(upg, fh, separators=(",", ":"))
    # ---- vcf2scoary (scoary/vcf2scoary.py) on its example and on a generated multi-allelic VCF
    import importlib
    conv = importlib.import_module("scoary.vcf2scoary")
    vdir = os.path.join(HERE, "vcf")
    os.makedirs(vdir, exist_ok=True)
    shutil.copyfile(os.path.join(EX, "Example.vcf"), os.path.join(vdir, "Example.vcf"))
    vr = random.Random(5)
    vl = ["##fileformat=VCFv4.2", '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
          '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Depth">',
          '##INFO=<ID=TYPE,Number=A,Type=String,Description="The type of allele.">']
    samples = ["S%02d" % i for i in range(37)]
    vl.append("\t".join(["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + samples))
    for k in range(300):
        nalt = vr.choice([1, 1, 1, 2, 3])
        alts = ",".join(vr.sample("ACGT", nalt))
        typ = vr.choice(["snp", "ins", "del", "complex"])
        cells = []
        for _ in samples:
            g = vr.choice(["0", "0", "1", ".", "2" if nalt > 1 else "1", "3" if nalt > 2 else "0"])
            cells.append(g + ":" + str(vr.randint(1, 99)))
        vl.append("\t".join(["chr%d" % (k % 3), str(100 + k * 7), "id%d" % k if k % 4 else ".", "A", alts, "99", "PASS",
                             "TYPE=%s" % typ, "GT:DP"] + cells))
    with open(os.path.join(vdir, "generated.vcf"), "w") as fh:
        fh.write("\n".join(vl) + "\n")
    for vname, vtypes in (("Example", None), ("generated", None), ("generated", "snp,del")):
        outp = os.path.join(vdir, vname + ("_snp_del" if vtypes else "") + ".csv")
        old = sys.argv
        sys.argv = ["vcf2scoary", "--force", "--out", outp] + (["--types", vtypes] if vtypes else []) + \
                   [os.path.join(vdir, vname + ".vcf")]
        try:
            conv.main()
        except SystemExit:
            pass
        sys.argv = old
    # ---- the reference CLI on the converted tables (-s 11): golden for reading a VCF directly (-g x.vcf)
    with open(os.path.join(vdir, "generated_traits.csv"), "w") as fh:
        fh.write(",resistant,with_missing\n")
        for i, smp in enumerate(samples):
            fh.write("%s,%d,%s\n" % (smp, vr.randint(0, 1), "NA" if i % 9 == 4 else str(vr.randint(0, 1))))
    shutil.copyfile(os.path.join(EX, "ExampleVCFTrait.csv"), os.path.join(vdir, "ExampleVCFTrait.csv"))
    for vname, tname in (("generated", "generated_traits.csv"), ("Example", "ExampleVCFTrait.csv")):
        out = tempfile.mkdtemp(prefix="golden_vcfcli_")
        rc = ref_shim.run_cli(["-g", os.path.join(vdir, vname + ".csv"), "-t", os.path.join(vdir, tname), "-s", "11",
                               "-p", "1.0", "-c", "I", "-o", out, "--no-time"])
        assert rc in (0, None), rc
        for f in sorted(os.listdir(out)):
            if f.endswith(".results.csv"):
                store(os.path.join(out, f), os.path.join(vdir, "cli_" + vname, f), False)
        shutil.rmtree(out)
    # ---- the reference's own CI golden row
    with open(os.path.join(HERE, "tetrcg_first_row.json"), "w") as fh:
        json.dump({"source": "tests/test_scoary_output.py:12-14",
                   "row": ["TetRCG", "", "A fictitious gene known to cause resistance against tetracycline", 29, 8, 3,
                           60, 90.625, 88.2352941176, 72.5, 1.08621066108E-014, 6.45209132679E-011,
                           6.45209132679E-011, 25, 25, 1, 5.96046447754E-008, 1.54972076416E-006]}, fh)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
