"""Reference-rule mode at the bench shape with different labellings per round (SB_RULE_ROUND_LABELLINGS): one line each."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json; sys.path.insert(0, %(root)r)
import numpy as np, torch
from scoary_b200 import synth
from scoary_b200.engine import Engine
from scoary_b200.methods import early_stop_table
G, N, P, seed = 50000, 5000, %(P)d, 20260903
traits = synth.make_traits(N, 1, seed)
cache = "/tmp/sb_sweep_bits_{}_{}_{}.npy".format(G, N, seed)
import os
bits = np.load(cache) if os.path.exists(cache) else synth.make_genes_packed(G, N, seed, traits=traits)
if not os.path.exists(cache): np.save(cache, bits)
col = {n: j for j, n in enumerate(synth.isolate_names(N))}
e = Engine(0); e.set_genes(bits, N); e.set_trait_vector(0, traits[0]); e.set_tree_nested(0, synth.make_tree(N, seed), col)
rm = early_stop_table(P)
e.permute(0, P, seed=seed, early_stop=True, rmin=rm)
e.set_profiling(True); e.stats_reset()
import time; t0 = time.perf_counter()
pairs, r, nd = e.permute(0, P, seed=seed, early_stop=True, rmin=rm)
dt = time.perf_counter() - t0
st = e.stats()
print(json.dumps({"wall_ms": dt * 1e3, "ms_permute": st["ms_permute"], "launches": st["launches_permute"], "walks": st["tests_walks"],
                  "ms_reduce": st["ms_reduce"], "ms_shuffle": st["ms_shuffle"], "crc": int(r.sum()) + int(nd.sum())}))
'''
for P in (1000, 10000):
    for n in ("", "37", "74", "91", "60", "45"):
        env = dict(os.environ)
        if n: env["SB_RULE_ROUND_LABELLINGS"] = n
        res = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "P": P}], env=env, capture_output=True, text=True)
        print("P=%d labellings/round=%s" % (P, n or "plan"), res.stdout.strip().splitlines()[-1] if res.returncode == 0 else res.stderr[-300:])
