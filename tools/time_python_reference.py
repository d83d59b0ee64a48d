#!/usr/bin/env python
"""Time the UNMODIFIED Python reference (shimmed, oracle/ref_shim.py) on small samples of the bench workloads.

Runs in the BUILD container only -- /root/reference does not travel to the GPU box, so bench.py cannot time the Python
reference in its own run; it quotes the figures this script writes to profiles/python_reference_timing.json as
`cpu_baseline.python_reference` and says where they were measured.  What is timed:

  Setup_results            scoary/methods.py:757-928 on G_s genes x N isolates (single process by design)
  ConvertUPGMAtoPhyloTree  scoary/methods.py:1386-1402 (one tree walk) at N isolates
  Permute                  scoary/methods.py:1314-1369 with the early stop disabled (binom.cdf stubbed to 0: exhaustive
                           mode, as the bench's headline), P_s permutations

  python tools/time_python_reference.py            # ~2 minutes on one core
"""
import json
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import ref_shim  # noqa: E402
from scoary_b200 import synth  # noqa: E402


def dicts(G, N, seed):
    traits = synth.make_traits(N, 1, seed)
    m = synth.unpack_rows(synth.make_genes_packed(G, N, seed, traits=traits), N)
    names = synth.isolate_names(N)
    genedic = {}
    for g in range(G):
        d = {names[j]: int(m[g, j]) for j in range(N)}
        d["Non-unique Gene name"], d["Annotation"] = "", "synthetic"      # scoary/methods.py:472-475
        genedic["g%07d" % g] = d
    traitsdic = {"T": {names[j]: str(int(traits[0][j])) for j in range(N)}}
    return genedic, traitsdic, names, m, traits[0]


def main():
    M = ref_shim.load()
    out = {"where": "build container (no GPU box: /root/reference does not travel)", "host": platform.processor() or platform.machine(),
           "cpu_count": os.cpu_count(), "python": platform.python_version(), "rows": []}
    import scipy
    out["scipy"] = scipy.__version__
    for N, G_s, P_s, n_walk in ((1000, 300, 20, 10), (5000, 40, 10, 2)):
        genedic, traitsdic, names, m, t = dicts(G_s, N, 20260903)
        t0 = time.perf_counter()
        res = M.Setup_results(genedic, traitsdic, False)
        t_setup = time.perf_counter() - t0
        tested = len(res["Results"]["T"])
        tree = synth.make_tree(N, 20260903)
        gtc = res["Gene_trait_combinations"]["T"]
        genes = list(res["Results"]["T"])[:n_walk]
        t0 = time.perf_counter()
        for g in genes:
            M.ConvertUPGMAtoPhyloTree(tree, gtc[g])
        t_walk = (time.perf_counter() - t0) / len(genes)
        cdf = M.ss.binom.cdf
        M.ss.binom.cdf = lambda *a, **k: 0.0          # never stop early: exhaustive mode
        try:
            t0 = time.perf_counter()
            M.Permute(tree, dict(gtc[genes[0]]), P_s, {})
            t_perm = time.perf_counter() - t0
        finally:
            M.ss.binom.cdf = cdf
        row = {"isolates": N, "genes_sampled": G_s, "genes_tested": tested,
               "Setup_results_tests_per_s_per_core": tested / t_setup,
               "walk_ms": t_walk * 1e3, "walks_per_s_per_core": 1.0 / t_walk,
               "Permute_permutations": P_s, "Permute_tests_per_s_per_core": (P_s + 1) / t_perm}
        # tests/s of the whole path at this isolate count and P = 1000 permutations per gene (per core; the reference's
        # own parallelism is one process per --threads over genes, i.e. at best x cores)
        per_gene = t_setup / tested + 1001 * (t_perm / (P_s + 1))
        row["tests_per_s_per_core_at_P1000"] = 1001 / per_gene
        out["rows"].append(row)
        print(json.dumps(row))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "python_reference_timing.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
