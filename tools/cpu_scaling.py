"""Thread scaling of the oracle's C port on this host (diagnostic for cpu_baseline)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from scoary_b200 import synth
N, G, P = 5000, 1024, 8
traits = synth.make_traits(N, 1, 1); bits = synth.make_genes_packed(G, N, 1, traits=traits)
nested = synth.make_tree(N, 1); l, r, names = O.flatten_tree(nested)
col = {n: j for j, n in enumerate(synth.isolate_names(N))}; cols = np.array([col[n] for n in names])
g = np.ascontiguousarray(synth.unpack_rows(bits, N)[:, cols]); lab = traits[0][cols].astype(np.uint8)
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for th in (1, 8, 32, 64, 128):
    if th > (os.cpu_count() or 1): break
    O.set_num_threads(th)
    O.permute(l, r, g[:th * 2], lab, P=1, seed=1)
    t0 = time.time(); O.permute(l, r, g, lab, P=P, seed=1); dt = time.time() - t0
    print("threads %3d: %.0f walks/s (%.0f per thread)" % (th, G * (P + 1) / dt, G * (P + 1) / dt / th))
