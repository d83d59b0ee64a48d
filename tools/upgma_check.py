#!/usr/bin/env python
"""sb_upgma with four kernels per merge step against SB_UPGMA_FUSED=1 (two per step): the same merges, and the time
of each (run on the GPU box).  Small cases are also checked against the oracle's restatement of the reference."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import oracle as O  # noqa: E402
from scoary_b200 import engine as eng, synth  # noqa: E402


def timed(e, fused):
    os.environ["SB_UPGMA_FUSED"] = "1" if fused else "0"
    best, out = 1e30, None
    for _ in range(2):
        t0 = time.perf_counter()
        out = e.upgma()
        best = min(best, (time.perf_counter() - t0) * 1e3)
    return out, best


def main():
    res = {}
    with eng.Engine(0) as e:
        rng = np.random.default_rng(11)
        for G, N in ((300, 17), (64, 2), (2000, 100), (5000, 333)):          # vs the oracle, ties included
            m = (rng.random((G, N)) < rng.uniform(0.05, 0.95, size=(G, 1))).astype(np.uint8)
            if N > 9:
                m[:, 5] = m[:, 3]
                m[:, 9] = m[:, 3]
            e.set_genes(eng.pack_rows(m), N)
            want = O.upgma_merges(m)
            for fused in (False, True):
                got, _ = timed(e, fused)
                res["oracle_%d_%d_%s" % (G, N, "fused" if fused else "plain")] = bool(np.array_equal(got, want))
        m = (rng.random((150, 1500)) < 0.5).astype(np.uint8)                  # few genes: many tied distances
        e.set_genes(eng.pack_rows(m), 1500)
        a, ta = timed(e, False)
        b, tb = timed(e, True)
        res["ties_1500"] = {"equal": bool(np.array_equal(a, b)), "ms_plain": ta, "ms_fused": tb}
        G, N, seed = 50000, 5000, 20260903
        traits = synth.make_traits(N, 1, seed)
        e.set_genes(synth.make_genes_packed(G, N, seed, traits=traits), N)
        a, ta = timed(e, False)
        b, tb = timed(e, True)
        res["c3_5000"] = {"equal": bool(np.array_equal(a, b)), "ms_plain": ta, "ms_fused": tb}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
