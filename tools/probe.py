"""Quick device probe: int32 peak + kernel timings at a given shape (not a bench)."""
import argparse, json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scoary_b200 import synth, tree as treemod
from scoary_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--genes", type=int, default=50000)
ap.add_argument("--isolates", type=int, default=5000)
ap.add_argument("--perms", type=int, default=100)
ap.add_argument("--seed", type=int, default=20260903)
a = ap.parse_args()
e = Engine(0)
print("int32 add+max peak (VIADDMNMX): %.3e ops/s" % e.int32_peak(8192))
t0 = time.time()
traits = synth.make_traits(a.isolates, 1, a.seed)
bits = synth.make_genes_packed(a.genes, a.isolates, a.seed, traits=traits)
nested = synth.make_tree(a.isolates, a.seed)
print("synth %.1fs" % (time.time() - t0))
names = synth.isolate_names(a.isolates)
col = {n: j for j, n in enumerate(names)}
e.set_profiling(True)
e.set_genes(bits, a.isolates)
e.set_trait_vector(0, traits[0])
e.set_tree_nested(0, nested, col)
for rep in range(2):
    e.stats_reset()
    t0 = time.time(); counts, p, _ = e.contingency_fisher(0); t1 = time.time()
    pairs, r, nd = e.permute(0, a.perms, seed=1); t2 = time.time()
    st = e.stats()
    print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}))
    walks = a.genes * a.perms
    print("rep %d: fisher wall %.3fs, permute wall %.3fs; K5 %.1f ms -> %.3e walks/s" % (
        rep, t1 - t0, t2 - t1, st["ms_permute"], walks / (st["ms_permute"] * 1e-3)))
print("pairs head", pairs[:3].tolist(), "r head", r[:12].tolist(), "p head", p[:3].tolist())

# ---- e2e call breakdown (host buffers)
import time as _t
def tm(label, fn):
    t0 = _t.perf_counter(); r = fn(); e.synchronize(); print("  %-28s %8.2f ms" % (label, (_t.perf_counter() - t0) * 1e3)); return r
from scoary_b200 import tree as treemod
left, right, leaf_names = treemod.flatten(nested)
leaf_cols = np.asarray([col[n] for n in leaf_names], dtype=np.int32)
for rep in range(2):
    print("e2e breakdown rep", rep)
    tm("set_genes", lambda: e.set_genes(bits, a.isolates))
    tm("set_trait_vector", lambda: e.set_trait_vector(0, traits[0]))
    tm("set_tree", lambda: e.set_tree(0, left, right, leaf_cols))
    tm("contingency_fisher", lambda: e.contingency_fisher(0))
    tm("permute(P=%d)" % a.perms, lambda: e.permute(0, a.perms, seed=1))
    from scoary_b200.methods import early_stop_table
    rm = early_stop_table(a.perms)
    e.stats_reset()
    out = tm("permute early_stop", lambda: e.permute(0, a.perms, seed=1, early_stop=True, rmin=rm))
    print("   walks executed", e.stats()["tests_walks"], "of", a.genes * (a.perms + 1), "; stopped early:", int((out[2] < a.perms).sum()))

# ---- GPU tree construction timing
e.set_genes(bits, a.isolates)
for rep in range(2):
    merges = tm("upgma (N=%d, G=%d)" % (a.isolates, a.genes), lambda: e.upgma())
print("first merges", merges[:3].tolist())
