#!/usr/bin/env python
"""Build and time kernel variants side by side (one `gpurun` call for a whole sweep).

  python tools/sweep_variants.py build      # here (no GPU): nvcc builds variants/<name>.so for every entry of VARIANTS
  python tools/sweep_variants.py run        # on the GPU box: times K5 / Fisher for each library on the C3 shape and
                                            # checks that pair counts, hit counts and Fisher p equal the product library's

The product library is variant "base".  Variants differ only in -D switches the kernels already honour
(threads per block, genes per thread, resident blocks, Fisher block size); results must be identical."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "variants")
# name: compiler switches, or (git revision, switches) to build the kernels as they were at that revision
VARIANTS = {
    # round-2 sessions (profiles/r2_sweep.txt).  r2a: prmt 1.03x, minblocks6 1.03x, everything else <= 1.00x (two
    # labellings in lockstep 0.88-0.97x despite -20 % instructions: 12 warps per SM do not hide the ALU latency).
    # r2c (whole-wave launch planning in place, PRMT masks the default): 192 threads x 4 blocks per SM 1.09x,
    # 128 x 7 1.05x, 256 x 3 1.05x, 128 x 6 1.04x, without PRMT 0.97x.  r2d (192 x 4 the default): 192 x 5 0.97x,
    # 160 x 6 0.97x, 224 x 4 0.93x, 384 x 2 1.01x, 6 genes per thread 0.77x.  Last check, after the prologue lost
    # 12 registers of live state:
    # r2i: after the interpreter lost its whole-field compares and 10 live registers (a16 doubles as the 32-bit
    # accumulator of a pair's first gene): the kernels of the previous commit, more resident blocks, two labellings
    "before_875ceed": ("875ceed", ""),
    # r2i, first pass: the new interpreter 1.055x over 875ceed; 192 x 5 1.01x, 160 x 6 1.01x, 256 x 3 1.02x; two
    # labellings at 128 registers (16 warps) 0.99 - 1.01x with 18 % fewer instructions, at 96 registers 0.57x.
    # Second pass: six genes per thread at 96 registers (12 % fewer instructions, 18 - 20 warps per SM) 0.86 - 0.91x.
    # Third pass (fused bare cherries, single pops, in-place merges): 1.11x over 875ceed; 256 x 3, 192 x 5 and
    # 128 x 6 each 1.02x over 192 x 4 -> 128 x 6 adopted.  Fourth pass: 128-bit stack chunks in place
    # r2j (128-bit stack chunks in place): 192 x 4 0.98x, 256 x 3 1.00x.  r2k: the padded leaf stream 1.023x (no
    # window-crossing copies of the handlers: instruction fetch stalls) -> adopted; padded at 192 x 4 1.00x
    "crossing": "-DSB_WALK_PADDED=0",
}


def build():
    import tempfile
    os.makedirs(VDIR, exist_ok=True)
    for name, spec in VARIANTS.items():
        out = os.path.join(VDIR, name + ".so")
        rev, flags = spec if isinstance(spec, tuple) else (None, spec)
        with tempfile.TemporaryDirectory() as tmp:
            src = os.path.join(ROOT, "scoary_b200", "csrc")
            if rev:                                   # the sources of that revision, built out of tree
                tar = subprocess.run(["git", "-C", ROOT, "archive", rev, "scoary_b200/csrc", "include"], check=True,
                                     capture_output=True).stdout
                subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
                src = os.path.join(tmp, "scoary_b200", "csrc")
            subprocess.run(["make", "-B", "-C", src, "OUT=" + out, "EXTRA=" + flags], check=True, stdout=subprocess.DEVNULL)
        print("built", out)


CHILD = r'''
import json, sys, os, zlib
sys.path.insert(0, %(root)r)
import numpy as np
from scoary_b200 import synth
from scoary_b200.engine import Engine
G, N, P, seed = 50000, 5000, 360, 20260903
traits = synth.make_traits(N, 1, seed)
cache = "/tmp/sb_sweep_bits_{}_{}_{}.npy".format(G, N, seed)          # the same matrix for every variant: generate once
if os.path.exists(cache):
    bits = np.load(cache)
else:
    bits = synth.make_genes_packed(G, N, seed, traits=traits); np.save(cache, bits)
col = {n: j for j, n in enumerate(synth.isolate_names(N))}
e = Engine(0); e.set_profiling(True); e.set_genes(bits, N); e.set_trait_vector(0, traits[0]); e.set_tree_nested(0, synth.make_tree(N, seed), col)
best = {}
for rep in range(3):
    e.stats_reset(); c, p, _ = e.contingency_fisher(0); pairs, r, nd = e.permute(0, P, seed=1); st = e.stats()
    for k in ("ms_fisher", "ms_permute", "ms_walk"): best[k] = min(best.get(k, 1e30), st[k])
crc = zlib.crc32(pairs.tobytes() + r.tobytes() + c.tobytes())
np.save(os.environ["SB_SWEEP_P"], p)
print(json.dumps({"walks_per_s": G * P / (best["ms_permute"] * 1e-3), "crc": crc, **best}))
'''


def run():
    os.makedirs(VDIR, exist_ok=True)
    libs = [("base", os.path.join(ROOT, "scoary_b200", "libscoary_b200.so"))]
    libs += [(n, os.path.join(VDIR, n + ".so")) for n in VARIANTS if os.path.exists(os.path.join(VDIR, n + ".so"))]
    import numpy as np
    base = base_p = None
    for name, path in libs:
        pfile = os.path.join(VDIR, name + ".p.npy")
        env = dict(os.environ, SCOARY_B200_LIB=path, SB_SWEEP_P=pfile)
        res = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env, capture_output=True, text=True)
        if res.returncode != 0:
            print("%-16s FAILED: %s" % (name, res.stderr.strip().splitlines()[-1] if res.stderr.strip() else "?"))
            continue
        d = json.loads(res.stdout.strip().splitlines()[-1])
        p = np.load(pfile)
        if base is None:
            base, base_p = d, p
        ok = base_p > 1e-290
        perr = float(np.max(np.abs(p[ok] - base_p[ok]) / base_p[ok]))      # a new Fisher kernel may differ in the last bits
        print("%-16s K5 %.3e walks/s (%.2fx)  fisher %.3f ms (%.2fx)  integers %s  p max rel diff %.1e" % (
            name, d["walks_per_s"], d["walks_per_s"] / base["walks_per_s"], d["ms_fisher"], base["ms_fisher"] / d["ms_fisher"],
            "identical" if d["crc"] == base["crc"] else "DIFFER", perr))


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1] if len(sys.argv) > 1 else "build"]()
