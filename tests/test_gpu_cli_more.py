"""More drop-in checks on the GPU, added after the last GPU session of round 1 (first run: the round-end
GPU tier; conftest.py orders this file after the earlier parity tests): a VCF read directly, the edge cases
the differential fuzzer found, and the N > 1 command line (skipped on a one-GPU box)."""
import os

import pytest

from test_gpu_cli import GOLD, INT_COLS, _compare, _rows, _run, inputs  # noqa: F401  (helpers + fixture)

pytestmark = pytest.mark.gpu


def test_cli_two_gpus_matches_reference_results(inputs):
    """The N > 1 flow of the CLI on real GPUs (torchrun, NCCL).  Needs two GPUs: skipped on a one-GPU box
    (the same flow runs under gloo in tests/test_distributed_cpu.py)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29900 + os.getpid() % 90), "-m", "scoary_b200.methods", "-g", inputs["g"],
           "-t", inputs["t"], "-o", inputs["out"], "--no-time", "-p", "1.0", "-c", "I", "-e", "50"]
    res = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    for trait in ("Tetracycline_resistance", "Bogus_trait"):
        got = os.path.join(inputs["out"], trait + ".results.csv")
        hdr, rows = _rows(got)
        ghdr, gold = _rows(os.path.join(GOLD, "all", trait + ".results.csv.gz"))
        assert hdr[:len(ghdr)] == ghdr and hdr[-1] == "Empirical_p" and len(rows) == len(gold)
        col = {h: i for i, h in enumerate(hdr)}
        by_gene = {r[0]: r for r in rows}
        for g in gold:
            for c in INT_COLS:
                assert by_gene[g[0]][col[c]] == g[col[c]], (trait, g[0], c)


def test_cli_reads_a_vcf_directly(tmp_path):
    """`-g x.vcf` on the GPU against the reference CLI run on the vcf2scoary-converted table (-s 11)."""
    vdir = os.path.join(GOLD, "vcf")
    out = str(tmp_path / "out")
    _run(["-g", os.path.join(vdir, "generated.vcf"), "-t", os.path.join(vdir, "generated_traits.csv"), "-p", "1.0", "-c",
          "I", "-o", out, "--no-time"])
    for t in ("resistant", "with_missing"):
        _compare(os.path.join(out, t + ".results.csv"), os.path.join(vdir, "cli_generated", t + ".results.csv"),
                 key_cols=3)


def test_cli_edge_cases(tmp_path):
    """tests/golden/edge on the GPU: tied distances in sb_upgma (the reference's None cluster), an all-zero trait,
    genes without contrasting pairs, missing values."""
    edir = os.path.join(GOLD, "edge")
    out = str(tmp_path / "out")
    _run(["-g", os.path.join(edir, "genes.csv"), "-t", os.path.join(edir, "traits.csv"), "-p", "1.0", "-c", "I", "-u",
          "-o", out, "--no-time"])
    for f in ("perfect", "allzero", "withNA", "mixed"):
        _compare(os.path.join(out, f + ".results.csv"), os.path.join(edir, f + ".results.csv"))
        hdr, rows = _rows(os.path.join(out, f + ".results.csv"))
        _, gold = _rows(os.path.join(edir, f + ".results.csv"))
        for c in ("Odds_ratio", "Best_pairwise_comp_p", "Worst_pairwise_comp_p"):      # nan / inf cells as text
            assert [r[hdr.index(c)] for r in rows] == [r[hdr.index(c)] for r in gold], (f, c)
    with open(os.path.join(out, "Tree.nwk")) as a, open(os.path.join(edir, "Tree.nwk")) as b:
        assert a.read() == b.read()
