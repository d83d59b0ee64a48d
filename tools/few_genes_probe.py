"""Few genes x many permutations (what the CLI has left after decideifbreak): K5 with threads = genes, threads =
labellings and the automatic choice, at N = 5 000 isolates.  One JSON line per (genes, permutations, mode)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scoary_b200 import synth
from scoary_b200.engine import Engine
from scoary_b200.methods import early_stop_table

N, seed = 5000, 20260903
traits = synth.make_traits(N, 1, seed)
bits = synth.make_genes_packed(4096, N, seed, traits=traits)
col = {n: j for j, n in enumerate(synth.isolate_names(N))}
e = Engine(0)
e.set_genes(bits, N); e.set_trait_vector(0, traits[0]); e.set_tree_nested(0, synth.make_tree(N, seed), col)
e.set_profiling(True)
ref = {}
for S, P in ((64, 10000), (16, 10000), (1, 10000), (256, 10000), (64, 1000), (2048, 10000)):
    idx = np.arange(S, dtype=np.int64)
    for es in (False, True):
        for mode, name in ((1, "threads=genes"), (2, "threads=labellings"), (0, "auto")):
            e.set_permute_mode(mode)
            rm = early_stop_table(P) if es else None
            e.permute(0, P, seed=seed, gene_idx=idx, early_stop=es, rmin=rm)
            e.stats_reset()
            t0 = time.perf_counter()
            pairs, r, nd = e.permute(0, P, seed=seed, gene_idx=idx, early_stop=es, rmin=rm)
            wall = time.perf_counter() - t0
            st = e.stats()
            key = (S, P, es)
            crc = int(r.sum()) * 1000003 + int(nd.sum())
            ref.setdefault(key, crc)
            print(json.dumps({"genes": S, "permutations": P, "early_stop": es, "mode": name, "wall_ms": round(wall * 1e3, 3),
                              "ms_permute": round(st["ms_permute"], 3), "walks": int(st["tests_walks"]),
                              "walks_per_s_kernel": round((st["tests_walks"] - S) / (st["ms_permute"] * 1e-3)) if st["ms_permute"] else None,
                              "transposed": int(st["calls_transposed"]), "same_result": crc == ref[key]}))
e.set_permute_mode(0)
